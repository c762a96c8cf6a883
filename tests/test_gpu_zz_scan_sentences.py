"""GPU parity tests for aggregation 'scan-sentences' of AlignmentContrastiveLoss (alad/loss.py:136-149) through
the drop-in class: golden vectors of the unmodified reference (tests/golden/scan_sentences.npz) and the oracle on
seeded inputs.  Tolerances: scores <= 1e-4 relative in fp32 mode, <= 1e-2 absolute in bf16 mode; gradients
<= 1e-3 of the largest gradient magnitude in fp32 mode (3e-2 in bf16 mode).

The reference's gradient of this mode is NaN as soon as an image of the batch has masked regions (softmax over an
all -inf row, loss.py:139-140); it is compared where it is finite, and the oracle's gradient (masked regions get
none) everywhere else."""
import numpy as np
import pytest
import torch

from conftest import assert_scores_close, load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu

BF16_ATOL = 1e-2
FP32_RTOL = 1e-4


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


def crit(precision, **kw):
    from aladin_b200 import loss as L
    c = L.AlignmentContrastiveLoss(aggregation="scan-sentences", **kw)
    c.precision = precision
    return c


def assert_grad_close(got, ref, tol, what):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert np.isfinite(got).all(), what
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / scale
    assert err <= tol, f"{what}: max error {err:.3e} of the largest gradient entry > {tol:.1e}"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_scores_golden(precision):
    g = load_golden("scan_sentences")
    c = crit(precision)
    S = c(cu(g["a_im"]), cu(g["a_s"]), g["a_im_len"].tolist(), g["a_s_len"].tolist(), return_loss=False,
          return_similarity_mat=True)
    Sb = c(cu(g["b_im"]).permute(1, 0, 2), cu(g["b_s"]).permute(1, 0, 2), g["b_im_len"].tolist(), g["b_s_len"].tolist(),
           return_loss=False, return_similarity_mat=True)
    torch.cuda.synchronize()
    if precision == "fp32":
        assert_scores_close(S.cpu().numpy(), g["a_S"], FP32_RTOL, "case a")
        assert_scores_close(Sb.cpu().numpy(), g["b_S"], FP32_RTOL, "case b")
    else:
        assert np.abs(S.cpu().numpy() - g["a_S"]).max() <= BF16_ATOL
        assert np.abs(Sb.cpu().numpy() - g["b_S"]).max() <= BF16_ATOL


def test_degenerate_lengths_like_reference():
    """No valid region -> 0, no valid word -> NaN (softmax over an all -inf row), like the reference."""
    g = load_golden("scan_sentences")
    S = crit("fp32")(cu(g["c_im"]), cu(g["c_s"]), g["c_im_len"].tolist(), g["c_s_len"].tolist(), return_loss=False,
                     return_similarity_mat=True).cpu().numpy()
    ref = g["c_S"]
    assert np.array_equal(np.isnan(S), np.isnan(ref))
    assert np.all(S[1] == 0)
    ok = ~np.isnan(ref)
    assert_scores_close(S[ok], ref[ok], FP32_RTOL, "case c")


@pytest.mark.parametrize("key,mv", [("mv", True), ("sum", False)])
def test_hinge_loss_golden(key, mv):
    g = load_golden("scan_sentences")
    c = crit("fp32", margin=0.2, max_violation=mv)
    loss, S = c(cu(g["b_im"]).permute(1, 0, 2), cu(g["b_s"]).permute(1, 0, 2), g["b_im_len"].tolist(),
                g["b_s_len"].tolist(), return_loss=True, return_similarity_mat=True)
    np.testing.assert_allclose(loss.item(), g[f"b_loss_{key}"], rtol=1e-4)
    assert loss.dim() == 0 and S.shape == (7, 7)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
def test_gradients_golden_full_length_images(precision, tol):
    """Every image full length: the reference's autograd result is finite and must be reproduced."""
    g = load_golden("scan_sentences")
    il, sl = g["d_im_len"].tolist(), g["d_s_len"].tolist()
    im, s = cu(g["d_im"], True), cu(g["d_s"], True)
    c = crit(precision, margin=0.2, max_violation=True)
    S = c(im, s, il, sl, return_loss=False, return_similarity_mat=True)
    (S * cu(g["d_Gup"])).sum().backward()
    assert_grad_close(im.grad.cpu().numpy(), g["d_dim"], tol, "d im_set (dense upstream)")
    assert_grad_close(s.grad.cpu().numpy(), g["d_ds"], tol, "d s_seq (dense upstream)")
    if precision == "fp32":
        im, s = cu(g["d_im"], True), cu(g["d_s"], True)
        loss = c(im, s, il, sl)
        loss.backward()
        np.testing.assert_allclose(loss.item(), g["d_loss"], rtol=1e-4)
        assert_grad_close(im.grad.cpu().numpy(), g["d_dim_loss"], tol, "d im_set (hinge)")
        assert_grad_close(s.grad.cpu().numpy(), g["d_ds_loss"], tol, "d s_seq (hinge)")


def test_gradients_ragged_vs_oracle_and_reference_where_finite():
    g = load_golden("scan_sentences")
    il, sl = g["a_im_len"].tolist(), g["a_s_len"].tolist()
    im, s = cu(g["a_im"], True), cu(g["a_s"], True)
    S = crit("fp32")(im, s, il, sl, return_loss=False, return_similarity_mat=True)
    (S * cu(g["a_Gup"])).sum().backward()
    d_im, d_s = O.scan_backward(g["a_im"], g["a_s"], il, sl, g["a_Gup"])
    assert_grad_close(im.grad.cpu().numpy(), d_im, 1e-3, "d im_set vs oracle")
    assert_grad_close(s.grad.cpu().numpy(), d_s, 1e-3, "d s_seq vs oracle")
    full = [i for i, l in enumerate(il) if l == g["a_im"].shape[1]]
    assert_grad_close(im.grad.cpu().numpy()[full], g["a_dim"][full], 1e-3, "d im_set vs reference (full-length images)")
    # dropped slots (image slot 0, caption slots 0, -2, -1) and masked tokens get no gradient
    gi, gs = im.grad.cpu().numpy(), s.grad.cpu().numpy()
    assert np.all(gi[:, 0] == 0) and np.all(gs[:, 0] == 0) and np.all(gs[:, -2:] == 0)
    R, W, nr, nw = O.scored_extents(g["a_im"].shape, g["a_s"].shape, il, sl)
    for i in range(len(il)):
        assert np.all(gi[i, 1 + nr[i]:] == 0)
    for j in range(len(sl)):
        assert np.all(gs[j, 1 + nw[j]:] == 0)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_shape_vs_oracle_with_chunking(precision, monkeypatch):
    """BASELINE-shaped tokens (34 regions x 50 words), ragged, rectangular; forced image chunks must agree
    bit for bit with the single-block result."""
    from aladin_b200 import scan, synth
    im, s, il, sl = synth.raw_batch(31, 37, 29, 35, 53, 256, ragged=True, related=0.4)
    ref = O.scan_scores(im, s, il, sl)
    c = crit(precision)
    S1 = c(cu(im), cu(s), il, sl, return_loss=False, return_similarity_mat=True).cpu().numpy()
    monkeypatch.setattr(scan, "_CHUNK_BYTES", 34 * 29 * 50 * 4 * 5)           # 5 images per chunk
    S2 = c(cu(im), cu(s), il, sl, return_loss=False, return_similarity_mat=True).cpu().numpy()
    np.testing.assert_array_equal(S1, S2)
    if precision == "fp32":
        assert_scores_close(S1, ref, FP32_RTOL, "training shape")
    else:
        assert np.abs(S1 - ref).max() <= BF16_ATOL


def test_backward_chunked_equals_unchunked_and_oracle(monkeypatch):
    from aladin_b200 import scan, synth
    im, s, il, sl = synth.raw_batch(32, 11, 9, 35, 53, 128, ragged=True, related=0.4)
    r = np.random.RandomState(3)
    Gup = r.standard_normal((11, 9)).astype(np.float32)
    Gup[r.rand(11, 9) < 0.5] = 0.0                                            # skipped pairs
    grads = []
    for chunk in (None, 34 * 9 * 50 * 4 * 4):
        if chunk:
            monkeypatch.setattr(scan, "_CHUNK_BYTES", chunk)
        a, b = cu(im, True), cu(s, True)
        S = crit("fp32")(a, b, il, sl, return_loss=False, return_similarity_mat=True)
        (S * cu(Gup)).sum().backward()
        grads.append((a.grad.cpu().numpy(), b.grad.cpu().numpy()))
    d_im, d_s = O.scan_backward(im, s, il, sl, Gup)
    assert_grad_close(grads[0][0], d_im, 1e-3, "d im_set")
    assert_grad_close(grads[0][1], d_s, 1e-3, "d s_seq")
    assert_grad_close(grads[1][0], grads[0][0], 1e-5, "chunked d im_set")
    assert_grad_close(grads[1][1], grads[0][1], 1e-5, "chunked d s_seq")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-1)])
def test_hinge_backward_takes_the_pair_list_path(precision, tol, monkeypatch):
    """B = 32 with hardest negatives: <= 3B of the B^2 entries of dL/dS are non-zero -> pair-list backward
    (alad_scan_apply_pairs); must equal the oracle and the dense-GEMM backward.  bf16 mode is compared loosely with
    the fp64 oracle: at d = 64 the operand rounding (~2e-3 on a cosine) flips the relu mask of near-zero cosines,
    which moves single gradient entries by a few percent of the largest one (measured 6.4e-2); the pair-list and
    the dense backward see the same cosines and must agree tightly in either mode."""
    from aladin_b200 import scan, synth
    im, s, il, sl = synth.raw_batch(33, 32, 32, 9, 12, 64, ragged=True, related=0.5)
    calls = []
    orig = scan._cabi.check
    monkeypatch.setattr(scan._cabi, "check", lambda rc, what: (calls.append(what), orig(rc, what))[1])
    grads = {}
    for mode in ("sparse", "dense"):
        if mode == "dense":
            monkeypatch.setattr(scan, "SPARSE_FRACTION", 10 ** 9)
        a, b = cu(im, True), cu(s, True)
        c = crit(precision, margin=0.2, max_violation=True)
        loss, S = c(a, b, il, sl, return_loss=True, return_similarity_mat=True)
        loss.backward()
        grads[mode] = (a.grad.cpu().numpy(), b.grad.cpu().numpy(), S.detach().cpu().numpy())
    assert "alad_scan_apply_pairs" in calls
    G = O.triplet_grad(grads["sparse"][2], 0.2, True)
    assert 0 < np.count_nonzero(G) <= 3 * 32
    d_im, d_s = O.scan_backward(im, s, il, sl, G)
    assert_grad_close(grads["sparse"][0], d_im, tol, "d im_set (pair list) vs oracle")
    assert_grad_close(grads["sparse"][1], d_s, tol, "d s_seq (pair list) vs oracle")
    assert_grad_close(grads["sparse"][0], grads["dense"][0], 1e-4, "pair list vs dense d im_set")
    assert_grad_close(grads["sparse"][1], grads["dense"][1], 1e-4, "pair list vs dense d s_seq")


def test_eval_containers_through_i2t_callback():
    """Evaluation path: a closure over the drop-in with a non-MrSw aggregation is called per query like the
    reference does (alad/evaluation.py:199-210); ranks must equal the oracle ranking of the oracle scores."""
    from aladin_b200 import evaluation, synth
    images, captions, img_lens, cap_lens = synth.eval_containers(5, 60, 24, 64, max_regions=12, max_words=14)
    c = crit("fp32")

    def sim_fn(img, cap, img_len, cap_len):
        return c(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    evaluation.clear_cache()
    m, (ranks, top1) = evaluation.i2t(torch.from_numpy(images), torch.from_numpy(captions), img_lens, cap_lens,
                                      return_ranks=True, sim_function=sim_fn, cap_batches=2)
    S_ref = O.scan_scores(images[0::5], captions, img_lens[0::5], cap_lens)
    ri, t1 = O.i2t_ranks(S_ref)
    from conftest import assert_ranks_equal_up_to_ties
    gt = np.array([5 * i + int(np.argmax(S_ref[i, 5 * i:5 * i + 5])) for i in range(60)])
    assert_ranks_equal_up_to_ties(ranks, ri, S_ref, gt, FP32_RTOL, "i2t ranks")
    assert len(m) == 7
