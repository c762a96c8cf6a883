"""BASELINE.json configs at their full shapes:
  config 4  training step at B = 512, 34 regions x 50 words, d = 1024: scores of both precision modes, the three losses
            of ALADModel.forward_loss (alad/alad_model.py:371-428) and the alignment gradient against the oracle;
  config 2  COCO-1k retrieval at full shape (1000 x 5000, 34 x 50, d = 1024): complete rows / columns of S for sampled
            queries against the oracle, their ranks exactly;
  Recall@K  at COCO-1k and COCO-5k shape: the bf16-mode ranking against the fp32-mode ranking of the same gallery --
            every rank change must be explained by the 1e-2 score tolerance, R@1/5/10 are reported side by side."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


def test_training_step_512_vs_oracle():
    from aladin_b200 import alad_model as AM, synth
    B, d = 512, 1024
    im, s, il, cl = synth.raw_batch(11, B, B, 35, 53, d, related=0.6)
    r = np.random.RandomState(11)
    icls = r.standard_normal((B, d)).astype(np.float32)
    ccls = (0.7 * icls + r.standard_normal((B, d))).astype(np.float32)
    icls /= np.linalg.norm(icls, axis=1, keepdims=True)
    ccls /= np.linalg.norm(ccls, axis=1, keepdims=True)
    ref_S = O.mrsw_scores(im, s, il, cl, acc64=True)
    ref_M = O.dot_scores(icls, ccls)
    # fp32 mode: the three losses, scores, and the gradient of the alignment loss alone
    t_icls, t_ccls = cu(icls, True), cu(ccls, True)
    t_im, t_s = cu(im, True), cu(s, True)
    lm, la, ld, M, S = AM.train_losses(t_icls, t_ccls, t_im, t_s, il, cl, margin=0.2, max_violation=True, precision="fp32",
                                       precision_matching="fp32")
    assert_scores_close(S.cpu().numpy(), ref_S, 1e-4, "B=512 alignment scores (fp32 mode)")
    assert_scores_close(M.cpu().numpy(), ref_M, 1e-4, "B=512 matching scores", floor=CANCELLING_FLOOR)
    np.testing.assert_allclose(la.item(), O.triplet_loss(ref_S, 0.2, True), rtol=1e-4)
    np.testing.assert_allclose(lm.item(), O.triplet_loss(ref_M, 0.2, True), rtol=1e-4)
    np.testing.assert_allclose(ld.item(), O.listnet_loss(ref_S, ref_M), rtol=1e-4)
    la.backward()
    G = O.triplet_grad(S.cpu().numpy(), 0.2, True)          # hardest negatives as the CUDA path saw them
    d_im, d_s = O.mrsw_backward(im, s, il, cl, G)
    np.testing.assert_allclose(t_im.grad.cpu().numpy(), d_im, rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(t_s.grad.cpu().numpy(), d_s, rtol=2e-3, atol=2e-5)
    # bf16 mode (the mode the bench times): |dS| <= 1e-2
    with torch.no_grad():
        _, _, _, _, S16 = AM.train_losses(cu(icls), cu(ccls), cu(im), cu(s), il, cl, margin=0.2, max_violation=True,
                                          precision="bf16", precision_matching="fp32")
    assert np.abs(S16.cpu().numpy() - ref_S).max() <= 1e-2


@pytest.fixture(scope="module")
def coco1k():
    from aladin_b200 import retrieval, synth
    Ni, Nc = 1000, 5000
    images, captions, il, cl = synth.dense_gallery_device(Ni, Nc, 34, 50, 1024, alpha=0.05)
    out = {}
    for prec in ("fp32", "bf16"):
        S = retrieval.AlignmentGallery(images, captions, il, cl, n_images=Ni, precision=prec).scores()
        out[prec] = (S, retrieval.rank_both_directions(S, Ni, k=50))
    torch.cuda.synchronize()
    return dict(images=images, captions=captions, il=il, cl=cl, **out)


def test_coco1k_full_shape_sampled_queries_vs_oracle(coco1k):
    Ni, Nc = 1000, 5000
    S, (ri, t1, rt, t50) = coco1k["fp32"]
    r = np.random.RandomState(2)
    rows = np.sort(r.choice(Ni, 32, replace=False))
    cols = np.sort(r.choice(Nc, 120, replace=False))
    im_all, cap_all = coco1k["images"].cpu().numpy(), coco1k["captions"].cpu().numpy()
    # complete rows of the sampled query images: i2t ranks and top-1 exactly
    ref_rows = O.mrsw_scores(im_all[rows], cap_all, [35] * len(rows), coco1k["cl"], acc64=True)
    got_rows = S[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert_scores_close(got_rows, ref_rows, 1e-4, "COCO-1k rows (fp32 mode)")
    for k, i in enumerate(rows):
        order = np.argsort(ref_rows[k], kind="stable")[::-1]
        pos = np.empty(Nc, np.int64)
        pos[order] = np.arange(Nc)
        want = pos[5 * i:5 * i + 5].min()
        if ri[i] != want:                                   # only a tie inside the score tolerance may move a rank
            gt = ref_rows[k, 5 * i:5 * i + 5].max()
            assert abs(ri[i] - want) <= np.count_nonzero(np.abs(ref_rows[k] - gt) <= 2e-4 * np.abs(ref_rows).max())
    # complete columns of the sampled query captions: t2i ranks
    ref_cols = O.mrsw_scores(im_all, cap_all[cols], coco1k["il"], [53] * len(cols), acc64=True)
    got_cols = S[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    assert_scores_close(got_cols, ref_cols, 1e-4, "COCO-1k columns (fp32 mode)")
    n_same = 0
    for k, c in enumerate(cols):
        want = np.count_nonzero(ref_cols[:, k] > ref_cols[c // 5, k])
        n_same += int(rt[c] == want)
        if rt[c] != want:
            g = ref_cols[c // 5, k]
            assert abs(rt[c] - want) <= np.count_nonzero(np.abs(ref_cols[:, k] - g) <= 2e-4 * np.abs(ref_cols).max())
    assert n_same >= len(cols) - 2
    # bf16 mode on the same sample: |dS| <= 1e-2
    S16 = coco1k["bf16"][0]
    assert np.abs(S16[torch.from_numpy(rows).cuda()].cpu().numpy() - ref_rows).max() <= 1e-2


def _recall_identity(S32, out32, S16, out16, Ni, Nc, tag):
    """bf16-mode ranking vs fp32-mode ranking of one gallery.  Returns the report dict."""
    from aladin_b200 import retrieval
    ri32, _, rt32, _ = out32
    ri16, _, rt16, _ = out16
    dS = float((S16 - S32).abs().max().item())
    assert dS <= 1e-2, f"{tag}: bf16-mode scores differ from fp32-mode scores by {dS}"
    rep = {"gallery": tag, "max_abs_score_difference": dS,
           "i2t": {"fp32": retrieval.recall_tuple(ri32)[:3], "bf16": retrieval.recall_tuple(ri16)[:3],
                   "rank_changes": int(np.count_nonzero(ri32 != ri16)), "queries": int(Ni)},
           "t2i": {"fp32": retrieval.recall_tuple(rt32)[:3], "bf16": retrieval.recall_tuple(rt16)[:3],
                   "rank_changes": int(np.count_nonzero(rt32 != rt16)), "queries": int(Nc)}}
    # every rank change must be explained by the tolerance: between the two rankings the ground truth may only have
    # passed candidates whose fp32-mode score lies within 2 * dS of its own (both scores moved by at most dS)
    S32h = S32.cpu().numpy()
    eps = 2 * dS + 1e-6
    for i in np.nonzero(ri32 != ri16)[0]:
        row = S32h[i]
        gt = row[5 * i:5 * i + 5].max()
        near = np.count_nonzero(np.abs(row - gt) <= eps)
        # the best ground-truth caption itself may change: allow the near-tied band of every ground-truth caption
        near += sum(np.count_nonzero(np.abs(row - g) <= eps) for g in row[5 * i:5 * i + 5])
        assert abs(ri32[i] - ri16[i]) <= near, (tag, "i2t", int(i), ri32[i], ri16[i], near)
    for c in np.nonzero(rt32 != rt16)[0]:
        col = S32h[:, c]
        near = np.count_nonzero(np.abs(col - col[c // 5]) <= eps)
        assert abs(rt32[c] - rt16[c]) <= near, (tag, "t2i", int(c), rt32[c], rt16[c], near)
    # R@K: a query may enter / leave the top K only through such a near tie, so the recalls differ by at most the
    # number of rank changes; report both
    for d, n in (("i2t", Ni), ("t2i", Nc)):
        for a, b in zip(rep[d]["fp32"], rep[d]["bf16"]):
            assert abs(a - b) <= 100.0 * rep[d]["rank_changes"] / n + 1e-9
    return rep


def _save(rep):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"recall_identity_{rep['gallery']}.json"), "w") as f:
        json.dump(rep, f, indent=1)


def test_recall_at_k_bf16_mode_vs_fp32_mode_coco1k(coco1k):
    rep = _recall_identity(coco1k["fp32"][0], coco1k["fp32"][1], coco1k["bf16"][0], coco1k["bf16"][1], 1000, 5000, "coco1k")
    _save(rep)


def test_recall_at_k_bf16_mode_vs_fp32_mode_coco5k():
    from aladin_b200 import retrieval, synth
    Ni, Nc = 5000, 25000
    images, captions, il, cl = synth.dense_gallery_device(Ni, Nc, 34, 50, 1024)
    res = {}
    for prec in ("fp32", "bf16"):
        S = retrieval.AlignmentGallery(images, captions, il, cl, n_images=Ni, precision=prec).scores()
        res[prec] = (S, retrieval.rank_both_directions(S, Ni, k=50))
    torch.cuda.synchronize()
    rep = _recall_identity(res["fp32"][0], res["fp32"][1], res["bf16"][0], res["bf16"][1], Ni, Nc, "coco5k")
    _save(rep)
