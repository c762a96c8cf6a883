"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes ``tests/golden/*.npz`` (inputs + reference outputs, all small).  The GPU
box never runs this script; tests read the committed ``.npz`` files.

Reference entry points exercised (mesnico/ALADIN @ d1bd7bf):
  alad/loss.py:70-159   AlignmentContrastiveLoss  (all tensor pooling modes)
  alad/loss.py:162-186  ContrastiveLoss           (dot / cosine, max_violation on/off)
  alad/loss.py:359-447  DistillationLoss          (listnet)
  alad/evaluation.py:158-327  i2t / t2i           (alignment callback and slot-0 path)
  alad/recall_auxiliary.py:133-149  compute_recall

Optional arguments select single generators (e.g. `make_golden.py scan_sentences`).
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("ALAD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
# the reference calls .cuda() unconditionally in i2t/t2i (evaluation.py:179,202,267,291)
torch.Tensor.cuda = lambda self, *a, **k: self

import alad.loss as rloss  # noqa: E402
import alad.evaluation as reval  # noqa: E402
import alad.recall_auxiliary as rrec  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(4)


def rs(seed):
    return np.random.RandomState(seed)


def t(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), requires_grad=grad)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------------------------------
# 1. alignment scores, rectangular, ragged, every tensor pooling mode
# ---------------------------------------------------------------------------------------
def golden_alignment_scores():
    r = rs(11)
    Bi, Bc, S_im, S_s, d = 5, 7, 9, 12, 32
    im = r.standard_normal((Bi, S_im, d)).astype(np.float32) * 1.7
    s = r.standard_normal((Bc, S_s, d)).astype(np.float32) * 0.6
    im[2, 4] = 0.0                        # an all-zero token (normalize eps path)
    im_len = [9, 4, 1, 6, 9]              # full, ragged, zero scored regions (len-1 == 0)
    s_len = [12, 3, 7, 5, 12, 4, 9]       # full container, zero scored words (len-3 == 0), ragged
    out = {}
    for agg in ("sum", "mean", "MrSw", "MrAVGw", "symm", "MwSr"):
        crit = rloss.AlignmentContrastiveLoss(aggregation=agg)
        with torch.no_grad():
            S = crit(t(im), t(s), im_len, s_len, return_loss=False, return_similarity_mat=True)
        out["S_" + agg] = S.numpy()
    save("alignment_scores", im=im, s=s, im_len=np.array(im_len), s_len=np.array(s_len), **out)


# ---------------------------------------------------------------------------------------
# 2. alignment loss + gradients (square), max_violation on/off, permuted [S,B,d] inputs
# ---------------------------------------------------------------------------------------
def golden_alignment_loss():
    r = rs(12)
    B, S_im, S_s, d = 6, 8, 11, 48
    im_sbd = r.standard_normal((S_im, B, d)).astype(np.float32)
    s_sbd = r.standard_normal((S_s, B, d)).astype(np.float32)
    # make the diagonal pairs related so that the hinge has both active and inactive rows
    for b in range(B):
        s_sbd[1:6, b] += 1.5 * im_sbd[1:6, b]
    im_len = [8, 5, 8, 3, 6, 7]
    s_len = [11, 6, 9, 11, 5, 8]
    out = {}
    for mv in (True, False):
        im_t = t(im_sbd, True)
        s_t = t(s_sbd, True)
        crit = rloss.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=mv, aggregation="MrSw")
        loss, S = crit(im_t.permute(1, 0, 2), s_t.permute(1, 0, 2), im_len, s_len, return_similarity_mat=True)
        loss.backward()
        k = "mv" if mv else "sum"
        out[f"loss_{k}"] = loss.detach().numpy()
        out[f"S_{k}"] = S.detach().numpy()
        out[f"dim_{k}"] = im_t.grad.numpy()
        out[f"ds_{k}"] = s_t.grad.numpy()
    # dense upstream gradient on S (exercises the dense backward): L = sum(Gup * S)
    Gup = r.standard_normal((B, B)).astype(np.float32)
    im_t = t(im_sbd, True)
    s_t = t(s_sbd, True)
    crit = rloss.AlignmentContrastiveLoss(aggregation="MrSw")
    S = crit(im_t.permute(1, 0, 2), s_t.permute(1, 0, 2), im_len, s_len, return_loss=False, return_similarity_mat=True)
    (S * t(Gup)).sum().backward()
    out["Gup"] = Gup
    out["dim_dense"] = im_t.grad.numpy()
    out["ds_dense"] = s_t.grad.numpy()
    save("alignment_loss", im_sbd=im_sbd, s_sbd=s_sbd, im_len=np.array(im_len), s_len=np.array(s_len), **out)


# ---------------------------------------------------------------------------------------
# 3. matching head: ContrastiveLoss (dot, cosine) + gradients
# ---------------------------------------------------------------------------------------
def golden_matching():
    r = rs(13)
    B, d = 9, 40
    im = r.standard_normal((B, d)).astype(np.float32)
    s = (0.8 * im + r.standard_normal((B, d))).astype(np.float32)
    im /= np.linalg.norm(im, axis=1, keepdims=True)
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    out = {}
    for measure in ("dot", "cosine"):
        for mv in (True, False):
            im_t, s_t = t(im * (1.0 if measure == "dot" else 2.5), True), t(s, True)
            crit = rloss.ContrastiveLoss(margin=0.2, measure=measure, max_violation=mv)
            loss, S = crit(im_t, s_t, return_similarity_mat=True)
            loss.backward()
            k = f"{measure}_{'mv' if mv else 'sum'}"
            out[f"loss_{k}"] = loss.detach().numpy()
            out[f"S_{k}"] = S.detach().numpy()
            out[f"dim_{k}"] = im_t.grad.numpy()
            out[f"ds_{k}"] = s_t.grad.numpy()
    save("matching", im=im, s=s, **out)


# ---------------------------------------------------------------------------------------
# 4. triplet on raw matrices incl. ties / zero-violation rows; listnet distillation
# ---------------------------------------------------------------------------------------
def golden_triplet_listnet():
    r = rs(14)
    B = 10
    S = r.standard_normal((B, B)).astype(np.float32)
    S[np.arange(B), np.arange(B)] += 1.0
    S[3, :] = -5.0
    S[3, 3] = 5.0                         # row with no violation
    S[:, 7] = np.minimum(S[:, 7], -4.0)
    S[7, 7] = 4.0                         # column with no violation
    S[5, 1] = S[5, 2] = 3.0               # tie for the hardest negative of row 5
    out = {"S": S}
    crit = rloss.Contrastive(margin=0.2, measure="dot", max_violation=True)
    for mv in (True, False):
        crit.max_violation = mv
        S_t = t(S, True)
        loss = crit.compute_contrastive_loss(S_t)
        loss.backward()
        k = "mv" if mv else "sum"
        out[f"loss_{k}"] = loss.detach().numpy()
        out[f"G_{k}"] = S_t.grad.numpy()
    # listnet: teacher = alignment-like magnitudes (O(#words)), student = cosines
    T = (r.standard_normal((B, B)) * 1.5 + 4.0).astype(np.float32)
    M = np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32)
    T_t, M_t = t(T, True), t(M, True)
    dl = rloss.DistillationLoss(mode="listnet")
    loss = dl(T_t, M_t)
    loss.backward()
    out.update(T=T, M=M, listnet_loss=loss.detach().numpy(), listnet_dM=M_t.grad.numpy(),
               listnet_dT_is_none=np.array(T_t.grad is None))
    save("triplet_listnet", **out)


# ---------------------------------------------------------------------------------------
# 5. retrieval: i2t / t2i (alignment callback + slot-0 path) and compute_recall
# ---------------------------------------------------------------------------------------
def make_eval_containers(seed, Ni, S, d, max_regions, max_words):
    """[N,S,d] containers laid out like encode_data (evaluation.py:98-130): slot 0 =
    global vector, tokens from slot 1, zero padding, every image row repeated 5x."""
    r = rs(seed)
    N = 5 * Ni
    img_feat_len = r.randint(3, max_regions + 1, size=Ni)            # raw feat_len (incl. slot 0)
    cap_len = r.randint(4, max_words + 1, size=N)                    # [CLS] w.. [SEP]
    images = np.zeros((N, S, d), np.float32)
    captions = np.zeros((N, S, d), np.float32)
    base = r.standard_normal((Ni, S, d)).astype(np.float32)
    for i in range(Ni):
        L = img_feat_len[i]
        base[i, L:] = 0
        images[5 * i:5 * i + 5] = base[i]
    for c in range(N):
        L = cap_len[c]
        noise = r.standard_normal((L, d)).astype(np.float32)
        src = base[c // 5, 1 + r.randint(0, max(img_feat_len[c // 5] - 1, 1), size=L)]
        captions[c, :L] = noise + 0.55 * src
    img_lens = [int(img_feat_len[i // 5]) for i in range(N)]
    return images, captions, img_lens, [int(x) for x in cap_len]


def golden_retrieval():
    Ni, S, d = 60, 20, 24
    images, captions, img_lens, cap_lens = make_eval_containers(15, Ni, S, d, max_regions=12, max_words=15)
    crit = rloss.AlignmentContrastiveLoss(aggregation="MrSw")

    def sim_fn(img, cap, img_len, cap_len):
        with torch.no_grad():
            return crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    sink = io.StringIO()
    with contextlib.redirect_stderr(sink), contextlib.redirect_stdout(sink):
        m_i2t, (ranks_i2t, top1) = reval.i2t(ti, tc, img_lens, cap_lens, return_ranks=True,
                                              sim_function=sim_fn, cap_batches=5)
        m_t2i, (ranks_t2i, top50) = reval.t2i(ti, tc, img_lens, cap_lens, return_ranks=True,
                                               sim_function=sim_fn, im_batches=5)
        g_i2t, (granks_i2t, gtop1) = reval.i2t(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=None)
        g_t2i, (granks_t2i, gtop50) = reval.t2i(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=None)
        rec = rrec.compute_recall(ti[:, 0, :], tc[:, 0, :])
        S_full = sim_fn(ti[0::5], tc, img_lens[0::5], cap_lens).numpy()
    save("retrieval", images=images[0::5], captions=captions,        # image rows are 5x duplicates: store one each
         img_lens=np.array(img_lens), cap_lens=np.array(cap_lens),
         m_i2t=np.array(m_i2t, dtype=np.float64), ranks_i2t=ranks_i2t, top1=top1,
         m_t2i=np.array(m_t2i, dtype=np.float64), ranks_t2i=ranks_t2i, top50=top50,
         g_i2t=np.array(g_i2t, dtype=np.float64), granks_i2t=granks_i2t, gtop1=gtop1,
         g_t2i=np.array(g_t2i, dtype=np.float64), granks_t2i=granks_t2i, gtop50=gtop50,
         compute_recall=np.array(rec, dtype=np.float64), S_full=S_full)


# ---------------------------------------------------------------------------------------
# 6. training step composite: the three criterion calls of ALADModel.forward_loss
#    (alad/alad_model.py:377-405) and the weighting of ALADModel.forward (:445-453) for
#    loss-type 'alignment-distillation-matching', loss-weights [1, 1, 0.1]
#    (configs/alad-alignment-and-matching-triplet0.1-plus-distill.yaml:23-24)
# ---------------------------------------------------------------------------------------
def golden_train_step():
    r = rs(16)
    B, S_im, S_s, d = 12, 10, 13, 64
    img_set = r.standard_normal((S_im, B, d)).astype(np.float32)       # S x B x dim (alad_model.py:439)
    cap_seq = r.standard_normal((S_s, B, d)).astype(np.float32)
    for b in range(B):
        cap_seq[1:7, b] += 1.2 * img_set[1:7, b]
    img_cls = r.standard_normal((B, d)).astype(np.float32)
    cap_cls = (0.7 * img_cls + r.standard_normal((B, d))).astype(np.float32)
    img_cls /= np.linalg.norm(img_cls, axis=1, keepdims=True)
    cap_cls /= np.linalg.norm(cap_cls, axis=1, keepdims=True)
    img_len = [10, 7, 10, 4, 8, 9, 10, 5, 6, 10, 3, 8]
    cap_len = [13, 8, 11, 13, 6, 9, 12, 7, 13, 10, 5, 9]
    ts = [t(img_cls, True), t(cap_cls, True), t(img_set, True), t(cap_seq, True)]
    matching_criterion = rloss.ContrastiveLoss(margin=0.2, measure="dot", max_violation=True)
    alignment_criterion = rloss.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True, aggregation="MrSw")
    distillation_loss = rloss.DistillationLoss(mode="listnet")
    matching_loss, matching_mat = matching_criterion(ts[0], ts[1], return_similarity_mat=True)
    alignment_loss, teacher_scores = alignment_criterion(ts[2].permute(1, 0, 2), ts[3].permute(1, 0, 2), img_len, cap_len,
                                                         return_similarity_mat=True)
    dist = distillation_loss(teacher_scores, matching_mat)
    weights = {"alignment": 1.0, "distillation": 1.0, "matching": 0.1}
    loss = alignment_loss * weights["alignment"] + dist * weights["distillation"] + matching_loss * weights["matching"]
    loss.backward()
    save("train_step", img_cls=img_cls, cap_cls=cap_cls, img_set=img_set, cap_seq=cap_seq,
         img_len=np.array(img_len), cap_len=np.array(cap_len),
         matching_loss=matching_loss.detach().numpy(), alignment_loss=alignment_loss.detach().numpy(),
         distillation_loss=dist.detach().numpy(), loss=loss.detach().numpy(),
         matching_mat=matching_mat.detach().numpy(), teacher_scores=teacher_scores.detach().numpy(),
         d_img_cls=ts[0].grad.numpy(), d_cap_cls=ts[1].grad.numpy(), d_img_set=ts[2].grad.numpy(),
         d_cap_seq=ts[3].grad.numpy())


# ---------------------------------------------------------------------------------------
# 7. the remaining DistillationLoss modes (alad/loss.py:371-425) + gradients, order_sim,
#    and the gradients of the 'sum' / 'mean' pooling modes
# ---------------------------------------------------------------------------------------
def golden_distill_modes():
    r = rs(17)
    B = 11
    T = (r.standard_normal((B, B)) * 0.6 + 0.4).astype(np.float32)      # teacher: some entries below the 0.1 threshold
    T[np.arange(B), np.arange(B)] += 1.0
    M = np.clip(r.standard_normal((B, B)) * 0.35, -1, 1).astype(np.float32)
    M[np.arange(B), np.arange(B)] += 0.3
    out = {"T": T, "M": M}
    # mse (owns the wb parameter)
    dl = rloss.DistillationLoss(mode="mse")
    with torch.no_grad():
        dl.wb.copy_(torch.tensor([0.7, 0.25]))
    M_t = t(M, True)
    loss = dl(t(T), M_t)
    loss.backward()
    out.update(mse_wb=np.array([0.7, 0.25], np.float32), mse_loss=loss.detach().numpy(), mse_dM=M_t.grad.numpy(),
               mse_dwb=dl.wb.grad.numpy())
    # contrastive (hard negatives chosen by the teacher); the reference zeroes the teacher diagonal IN PLACE
    for margin in (0.2, 0.05):
        dl = rloss.DistillationLoss(mode="contrastive", margin=margin)
        M_t = t(M, True)
        T_t = t(T)
        loss = dl(T_t, M_t)
        loss.backward()
        k = f"contrastive_m{margin}"
        out.update({k + "_loss": loss.detach().numpy(), k + "_dM": M_t.grad.numpy(), k + "_T_after": T_t.numpy().copy()})
    # ordinal, default and non-default stride / threshold
    for (margin, thr, stride) in ((0.2, 0.1, 3), (0.1, 0.5, 1), (0.2, 100.0, 3)):
        dl = rloss.DistillationLoss(mode="ordinal", margin=margin, threshold=thr, stride=stride)
        M_t = t(M, True)
        loss = dl(t(T), M_t)
        k = f"ordinal_m{margin}_t{thr}_s{stride}"
        if torch.isfinite(loss):
            loss.backward()
            out[k + "_dM"] = M_t.grad.numpy()
        out[k + "_loss"] = loss.detach().numpy()
    # order_sim forward (alad/loss.py:20-26)
    im = np.abs(r.standard_normal((7, 20))).astype(np.float32)
    s = np.abs(r.standard_normal((9, 20))).astype(np.float32)
    out.update(order_im=im, order_s=s, order_S=rloss.order_sim(t(im), t(s)).numpy())
    save("distill_modes", **out)


def golden_pooled_grads():
    r = rs(18)
    Bi, Bc, S_im, S_s, d = 4, 5, 7, 9, 24
    im = r.standard_normal((Bi, S_im, d)).astype(np.float32) * 1.3
    s = r.standard_normal((Bc, S_s, d)).astype(np.float32) * 0.8
    im_len = [7, 3, 5, 7]
    s_len = [9, 5, 4, 9, 7]
    Gup = r.standard_normal((Bi, Bc)).astype(np.float32)
    out = dict(im=im, s=s, im_len=np.array(im_len), s_len=np.array(s_len), Gup=Gup)
    for agg in ("sum", "mean", "MrAVGw", "MwSr", "symm"):
        im_t, s_t = t(im, True), t(s, True)
        crit = rloss.AlignmentContrastiveLoss(aggregation=agg)
        S = crit(im_t, s_t, im_len, s_len, return_loss=False, return_similarity_mat=True)
        (S * t(Gup)).sum().backward()
        out.update({f"S_{agg}": S.detach().numpy(), f"dim_{agg}": im_t.grad.numpy(), f"ds_{agg}": s_t.grad.numpy()})
    save("pooled_grads", **out)


# ---------------------------------------------------------------------------------------
# 9. aggregation 'scan-sentences' (alad/loss.py:136-149): scores, hinge loss and gradients
# ---------------------------------------------------------------------------------------
def golden_scan_sentences():
    r = rs(23)
    out = {}
    # (a) rectangular, ragged, dense upstream gradient
    Bi, Bc, S_im, S_s, d = 5, 6, 8, 10, 24
    im = r.standard_normal((Bi, S_im, d)).astype(np.float32) * 1.3
    s = r.standard_normal((Bc, S_s, d)).astype(np.float32) * 0.8
    im_len = [8, 3, 5, 8, 2]
    s_len = [10, 5, 4, 10, 7, 6]
    Gup = r.standard_normal((Bi, Bc)).astype(np.float32)
    im_t, s_t = t(im, True), t(s, True)
    crit = rloss.AlignmentContrastiveLoss(aggregation="scan-sentences")
    S = crit(im_t, s_t, im_len, s_len, return_loss=False, return_similarity_mat=True)
    (S * t(Gup)).sum().backward()
    out.update(a_im=im, a_s=s, a_im_len=np.array(im_len), a_s_len=np.array(s_len), a_Gup=Gup, a_S=S.detach().numpy(),
               a_dim=im_t.grad.numpy(), a_ds=s_t.grad.numpy())
    # (b) square batch in the training layout ([S,B,d] permuted like alad_model.py:377-378), hinge loss with
    #     hardest negatives (sparse dL/dS) and with the plain sums
    B, S_im, S_s, d = 7, 9, 12, 32
    base = r.standard_normal((B, d)).astype(np.float32)
    im = (r.standard_normal((S_im, B, d)) + 0.8 * base[None]).astype(np.float32)
    s = (r.standard_normal((S_s, B, d)) + 0.8 * base[None]).astype(np.float32)
    im_len = [9, 9, 4, 6, 9, 2, 7]
    s_len = [12, 7, 12, 5, 9, 12, 4]
    out.update(b_im=im, b_s=s, b_im_len=np.array(im_len), b_s_len=np.array(s_len))
    for tag, mv in (("mv", True), ("sum", False)):
        im_t, s_t = t(im, True), t(s, True)
        crit = rloss.AlignmentContrastiveLoss(margin=0.2, max_violation=mv, aggregation="scan-sentences")
        loss, S = crit(im_t.permute(1, 0, 2), s_t.permute(1, 0, 2), im_len, s_len, return_loss=True,
                       return_similarity_mat=True)
        loss.backward()
        out.update({f"b_S": S.detach().numpy(), f"b_loss_{tag}": loss.detach().numpy(),
                    f"b_dim_{tag}": im_t.grad.numpy(), f"b_ds_{tag}": s_t.grad.numpy()})
    # (c) degenerate lengths: a caption without valid words (NaN column), an image without valid regions (0 row)
    Bi, Bc, S_im, S_s, d = 3, 4, 6, 8, 16
    im = r.standard_normal((Bi, S_im, d)).astype(np.float32)
    s = r.standard_normal((Bc, S_s, d)).astype(np.float32)
    im_len = [6, 1, 4]
    s_len = [8, 3, 5, 2]
    crit = rloss.AlignmentContrastiveLoss(aggregation="scan-sentences")
    with torch.no_grad():
        S = crit(t(im), t(s), im_len, s_len, return_loss=False, return_similarity_mat=True)
    out.update(c_im=im, c_s=s, c_im_len=np.array(im_len), c_s_len=np.array(s_len), c_S=S.numpy())
    # (d) every image full length (the only case in which the reference's gradient is finite: a masked region
    #     makes the softmax row all -inf, loss.py:139-140, and NaN * 0 reaches every scored word slot), ragged
    #     captions, dense upstream gradient, plus the hinge loss with hardest negatives
    B, S_im, S_s, d = 6, 7, 11, 40
    base = r.standard_normal((B, d)).astype(np.float32)
    im = (r.standard_normal((B, S_im, d)) + 0.1 * base[:, None]).astype(np.float32)
    s = (r.standard_normal((B, S_s, d)) + 0.1 * base[:, None]).astype(np.float32)
    im_len = [S_im] * B
    s_len = [11, 6, 4, 11, 8, 5]
    Gup = r.standard_normal((B, B)).astype(np.float32)
    im_t, s_t = t(im, True), t(s, True)
    crit = rloss.AlignmentContrastiveLoss(margin=0.2, max_violation=True, aggregation="scan-sentences")
    S = crit(im_t, s_t, im_len, s_len, return_loss=False, return_similarity_mat=True)
    (S * t(Gup)).sum().backward()
    out.update(d_im=im, d_s=s, d_im_len=np.array(im_len), d_s_len=np.array(s_len), d_Gup=Gup, d_S=S.detach().numpy(),
               d_dim=im_t.grad.numpy(), d_ds=s_t.grad.numpy())
    im_t, s_t = t(im, True), t(s, True)
    loss = crit(im_t, s_t, im_len, s_len)
    loss.backward()
    out.update(d_loss=loss.detach().numpy(), d_dim_loss=im_t.grad.numpy(), d_ds_loss=s_t.grad.numpy())
    assert all(np.isfinite(out[k]).all() for k in ("d_dim", "d_ds", "d_dim_loss", "d_ds_loss")) and float(loss) > 0
    save("scan_sentences", **out)


# ---------------------------------------------------------------------------------------
# 10. consumers of the ranking output: the ndcg_scorer hooks of i2t / t2i
#     (alad/evaluation.py:225-228,310-313) and recall_1k_5fold_test (alad/recall_auxiliary.py:90-130)
# ---------------------------------------------------------------------------------------
class RecordingScorer:
    """Stands in for evaluate_utils.dcg.DCG (which needs external relevance files): records what the reference
    hands to compute_ndcg; DCG itself only reads sorted_indexes[:rank], rank = 25 (dcg.py:19-20)."""

    def __init__(self):
        self.calls = []

    def compute_ndcg(self, npts, query_id, sorted_indexes, fold_index=0, retrieval='image'):
        self.calls.append((int(npts), int(query_id), np.asarray(sorted_indexes[:25]).astype(np.int64), int(fold_index), retrieval))
        return {'rougeL': 0.25, 'spice': 0.5}


def fold_embeddings(seed, n_rows, d):
    """Global vectors [n_rows, d] for the 5-fold test, regenerated from the seed by the tests (only the
    reference's OUTPUTS are stored): image rows repeated 5x, captions = image + noise."""
    r = rs(seed)
    base = r.standard_normal((n_rows // 5, d)).astype(np.float32)
    img = np.repeat(base, 5, axis=0)
    cap = (img + 1.5 * r.standard_normal((n_rows, d))).astype(np.float32)
    return img, cap


def golden_ranking_consumers():
    g = dict(np.load(os.path.join(HERE, "retrieval.npz")))
    images = np.repeat(g["images"], 5, axis=0)
    captions, img_lens, cap_lens = g["captions"], g["img_lens"].tolist(), g["cap_lens"].tolist()
    crit = rloss.AlignmentContrastiveLoss(aggregation="MrSw")

    def sim_fn(img, cap, img_len, cap_len):
        with torch.no_grad():
            return crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    out = {}
    sink = io.StringIO()
    with contextlib.redirect_stderr(sink), contextlib.redirect_stdout(sink):
        for tag, fn, kw in (("i2t", reval.i2t, dict(cap_batches=5)), ("t2i", reval.t2i, dict(im_batches=5))):
            sc = RecordingScorer()
            m = fn(ti, tc, img_lens, cap_lens, ndcg_scorer=sc, fold_index=2, sim_function=sim_fn, **kw)
            out[f"{tag}_metrics"] = np.array(m, dtype=np.float64)
            out[f"{tag}_query"] = np.array([c[1] for c in sc.calls])
            out[f"{tag}_order25"] = np.stack([c[2] for c in sc.calls])
            assert all(c[0] == 60 and c[3] == 2 and c[4] == ("sentence" if tag == "i2t" else "image") for c in sc.calls)
    # 5-fold recall: the unmodified function on 25000 rows (5 folds of 1000 images x 5000 captions), and the same
    # protocol on 5 folds of 50 images x 250 captions through the reference's recall_test (what the CPU tests use)
    img, cap = fold_embeddings(41, 25000, 8)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out["fold5000"] = np.array(rrec.recall_1k_5fold_test(torch.from_numpy(img), torch.from_numpy(cap)), dtype=np.float64)
    out["fold5000_stdout"] = np.array(buf.getvalue())
    img, cap = fold_embeddings(42, 1250, 8)
    with contextlib.redirect_stdout(io.StringIO()):
        folds = [rrec.recall_test(torch.from_numpy(img[k * 250:(k + 1) * 250]), torch.from_numpy(cap[k * 250:(k + 1) * 250]), None, None)
                 for k in range(5)]
    small = [float(np.mean([f[j] for f in folds])) for j in range(6)]
    out["fold250"] = np.array(small + [sum(small)], dtype=np.float64)
    save("ranking_consumers", **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (golden_alignment_scores, golden_alignment_loss, golden_matching, golden_triplet_listnet, golden_retrieval,
               golden_train_step, golden_distill_modes, golden_pooled_grads, golden_scan_sentences, golden_ranking_consumers):
        if not only or fn.__name__ in only or fn.__name__.replace("golden_", "") in only:
            fn()
