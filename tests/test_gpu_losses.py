"""GPU parity tests for the loss drop-ins (aladin_b200.loss) against the reference's golden
vectors and the oracle.  fp32 mode: scores <= 1e-4 relative; gradients compared with the
tolerances written at each assert."""
import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close, load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


@pytest.mark.parametrize("key,mv", [("mv", True), ("sum", False)])
def test_triplet_golden(key, mv):
    from aladin_b200 import loss as L
    g = load_golden("triplet_listnet")
    S = cu(g["S"], True)
    crit = L.Contrastive(margin=0.2, measure="dot", max_violation=mv)
    out = crit.compute_contrastive_loss(S)
    out.backward()
    np.testing.assert_allclose(out.item(), g[f"loss_{key}"], rtol=1e-6)
    np.testing.assert_array_equal(S.grad.cpu().numpy(), g[f"G_{key}"])       # integer-valued gradient: exact


@pytest.mark.parametrize("B", [1, 33, 512, 1500])
def test_triplet_vs_oracle(B):
    from aladin_b200 import loss as L
    r = np.random.RandomState(B)
    S = r.standard_normal((B, B)).astype(np.float32)
    for mv in (True, False):
        loss, G, ra, ca = L.triplet_fwd_bwd(cu(S), 0.2, mv)
        np.testing.assert_allclose(loss.item(), O.triplet_loss(S, 0.2, mv), rtol=2e-5)
        np.testing.assert_array_equal(G.cpu().numpy(), O.triplet_grad(S, 0.2, mv))


def test_listnet_golden():
    from aladin_b200 import loss as L
    g = load_golden("triplet_listnet")
    T, M = cu(g["T"], True), cu(g["M"], True)
    out = L.DistillationLoss(mode="listnet")(T, M)
    out.backward()
    np.testing.assert_allclose(out.item(), g["listnet_loss"], rtol=1e-5)
    np.testing.assert_allclose(M.grad.cpu().numpy(), g["listnet_dM"], rtol=2e-4, atol=1e-7)
    assert T.grad is None


@pytest.mark.parametrize("B", [7, 128, 512, 1100])
def test_listnet_vs_oracle(B):
    from aladin_b200 import loss as L
    r = np.random.RandomState(B)
    T = (r.standard_normal((B, B)) * 2 + 3).astype(np.float32)       # alignment-score magnitudes
    M = np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32)
    loss, dM = L.listnet_fwd_bwd(cu(T), cu(M))
    np.testing.assert_allclose(loss.item(), O.listnet_loss(T, M), rtol=2e-5)
    ref = O.listnet_grad(T, M)
    np.testing.assert_allclose(dM.cpu().numpy(), ref, rtol=1e-3, atol=1e-6 * np.abs(ref).max() + 1e-9)


def test_matching_golden():
    import aladin_b200
    from aladin_b200 import loss as L
    g = load_golden("matching")
    aladin_b200.set_precision("fp32")
    try:
        for key, mv in (("mv", True), ("sum", False)):
            im, s = cu(g["im"], True), cu(g["s"], True)
            loss, S = L.ContrastiveLoss(margin=0.2, measure="dot", max_violation=mv)(im, s, return_similarity_mat=True)
            loss.backward()
            k = f"dot_{key}"
            assert_scores_close(S.detach().cpu().numpy(), g[f"S_{k}"], 1e-4, k, floor=CANCELLING_FLOOR)
            np.testing.assert_allclose(loss.item(), g[f"loss_{k}"], rtol=1e-4)
            np.testing.assert_allclose(im.grad.cpu().numpy(), g[f"dim_{k}"], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(s.grad.cpu().numpy(), g[f"ds_{k}"], rtol=1e-4, atol=1e-5)
        # cosine measure, forward only
        _, S = L.ContrastiveLoss(margin=0.2, measure="cosine", max_violation=True)(cu(g["im"] * 2.5), cu(g["s"]), True)
        assert_scores_close(S.cpu().numpy(), g["S_cosine_mv"], 1e-4, "cosine", floor=CANCELLING_FLOOR)
    finally:
        aladin_b200.set_precision("bf16")


@pytest.mark.parametrize("key,mv", [("mv", True), ("sum", False)])
def test_alignment_loss_golden(key, mv):
    from aladin_b200 import loss as L
    g = load_golden("alignment_loss")
    im, s = cu(g["im_sbd"], True), cu(g["s_sbd"], True)
    crit = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=mv, aggregation="MrSw")
    crit.precision = "fp32"
    loss, S = crit(im.permute(1, 0, 2), s.permute(1, 0, 2), g["im_len"].tolist(), g["s_len"].tolist(),
                   return_similarity_mat=True)
    loss.backward()
    assert_scores_close(S.detach().cpu().numpy(), g[f"S_{key}"], 1e-4, key)
    np.testing.assert_allclose(loss.item(), g[f"loss_{key}"], rtol=1e-4)
    np.testing.assert_allclose(im.grad.cpu().numpy(), g[f"dim_{key}"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g[f"ds_{key}"], rtol=1e-3, atol=1e-5)


def test_alignment_dense_upstream_gradient_golden():
    from aladin_b200 import loss as L
    g = load_golden("alignment_loss")
    im, s = cu(g["im_sbd"], True), cu(g["s_sbd"], True)
    crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
    crit.precision = "fp32"
    S = crit(im.permute(1, 0, 2), s.permute(1, 0, 2), g["im_len"].tolist(), g["s_len"].tolist(), return_loss=False,
             return_similarity_mat=True)
    (S * cu(g["Gup"])).sum().backward()
    np.testing.assert_allclose(im.grad.cpu().numpy(), g["dim_dense"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["ds_dense"], rtol=1e-3, atol=1e-5)


def test_train_step_call_site_golden():
    """The three criterion calls of ALADModel.forward_loss + weighting (alad_model.py:377-405,445-453)."""
    import aladin_b200
    from aladin_b200 import loss as L
    g = load_golden("train_step")
    aladin_b200.set_precision("fp32")
    try:
        img_cls, cap_cls = cu(g["img_cls"], True), cu(g["cap_cls"], True)
        img_set, cap_seq = cu(g["img_set"], True), cu(g["cap_seq"], True)
        matching_criterion = L.ContrastiveLoss(margin=0.2, measure="dot", max_violation=True)
        alignment_criterion = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True, aggregation="MrSw")
        distillation_loss = L.DistillationLoss(mode="listnet")
        matching_loss, matching_mat = matching_criterion(img_cls, cap_cls, return_similarity_mat=True)
        alignment_loss, teacher = alignment_criterion(img_set.permute(1, 0, 2), cap_seq.permute(1, 0, 2),
                                                      g["img_len"].tolist(), g["cap_len"].tolist(), return_similarity_mat=True)
        dist = distillation_loss(teacher, matching_mat)
        loss = alignment_loss * 1.0 + dist * 1.0 + matching_loss * 0.1
        loss.backward()
        assert_scores_close(teacher.detach().cpu().numpy(), g["teacher_scores"], 1e-4, "teacher")
        assert_scores_close(matching_mat.detach().cpu().numpy(), g["matching_mat"], 1e-4, "matching", floor=CANCELLING_FLOOR)
        np.testing.assert_allclose([matching_loss.item(), alignment_loss.item(), dist.item(), loss.item()],
                                   [g["matching_loss"], g["alignment_loss"], g["distillation_loss"], g["loss"]], rtol=1e-4)
        np.testing.assert_allclose(img_cls.grad.cpu().numpy(), g["d_img_cls"], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(cap_cls.grad.cpu().numpy(), g["d_cap_cls"], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(img_set.grad.cpu().numpy(), g["d_img_set"], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(cap_seq.grad.cpu().numpy(), g["d_cap_seq"], rtol=1e-3, atol=1e-5)
    finally:
        aladin_b200.set_precision("bf16")


def test_training_batch_128_vs_oracle():
    """BASELINE config 1: B=128, 34 regions x 50 words (raw 35/53 slots), d=1024, ragged."""
    from aladin_b200 import loss as L, synth
    im, s, il, cl = synth.raw_batch(9, 128, 128, 35, 53, 1024, related=0.6)
    im_t, s_t = cu(im.transpose(1, 0, 2).copy(), True), cu(s.transpose(1, 0, 2).copy(), True)
    crit = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True, aggregation="MrSw")
    crit.precision = "fp32"
    loss, S = crit(im_t.permute(1, 0, 2), s_t.permute(1, 0, 2), il, cl, return_similarity_mat=True)
    loss.backward()
    ref = O.mrsw_scores(im, s, il, cl, acc64=True)
    assert_scores_close(S.detach().cpu().numpy(), ref, 1e-4, "B=128 scores")
    np.testing.assert_allclose(loss.item(), O.triplet_loss(ref, 0.2, True), rtol=1e-4)
    G = O.triplet_grad(S.detach().cpu().numpy(), 0.2, True)
    d_im, d_s = O.mrsw_backward(im, s, il, cl, G)
    np.testing.assert_allclose(im_t.grad.cpu().numpy().transpose(1, 0, 2), d_im, rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(s_t.grad.cpu().numpy().transpose(1, 0, 2), d_s, rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("agg", ["MwSr", "symm", "MrAVGw"])
def test_other_pooling_modes_gradients_vs_torch_autograd(agg):
    """Gradients of the non-default pooling modes against torch autograd on a plain fp32
    restatement of alad/loss.py:80-135 (the kernel under test is the CUDA backward)."""
    from aladin_b200 import loss as L, synth
    im, s, il, cl = synth.raw_batch(21, 7, 9, 8, 12, 48, related=0.7)
    cl = [max(c, 5) for c in cl]                                    # MrAVGw: keep #words >= 1 (no 0/0)
    im_t, s_t = cu(im, True), cu(s, True)
    crit = L.AlignmentContrastiveLoss(aggregation=agg)
    crit.precision = "fp32"
    S = crit(im_t, s_t, il, cl, return_loss=False, return_similarity_mat=True)
    Gup = cu(np.random.RandomState(5).standard_normal(S.shape))
    (S * Gup).sum().backward()
    # reference-shaped fp32 computation in torch (CPU, double) with autograd
    a = torch.tensor(im, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(s, dtype=torch.float64, requires_grad=True)
    an = torch.nn.functional.normalize(a, dim=2)[:, 1:]
    bn = torch.nn.functional.normalize(b, dim=2)[:, 1:-2]
    A = torch.einsum("ird,jwd->ijrw", an, bn)
    R, W = an.shape[1], bn.shape[1]
    rm = torch.arange(R)[None, :] >= torch.tensor([l - 1 for l in il])[:, None]
    wm = torch.arange(W)[None, :] >= torch.tensor([l - 3 for l in cl])[:, None]
    A = A.masked_fill(rm[:, None, :, None] | wm[None, :, None, :], 0.0)
    mrsw = A.max(2)[0].sum(2)
    mwsr = A.max(3)[0].sum(2)
    ref = {"MwSr": mwsr, "symm": mrsw + mwsr, "MrAVGw": mrsw / torch.tensor([l - 3 for l in cl], dtype=torch.float64)[None, :]}[agg]
    assert_scores_close(S.detach().cpu().numpy(), ref.detach().numpy(), 1e-4, agg)
    (ref * Gup.cpu().double()).sum().backward()
    np.testing.assert_allclose(im_t.grad.cpu().numpy(), a.grad.numpy(), rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(s_t.grad.cpu().numpy(), b.grad.numpy(), rtol=2e-3, atol=2e-5)


def _autograd_reference(im, s, il, cl, agg, Gup):
    """float64 torch restatement of alad/loss.py:80-135 with autograd: (S, d im, d s)."""
    a = torch.tensor(im, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(s, dtype=torch.float64, requires_grad=True)
    an = torch.nn.functional.normalize(a, dim=2)[:, 1:]
    bn = torch.nn.functional.normalize(b, dim=2)[:, 1:-2]
    A = torch.einsum("ird,jwd->ijrw", an, bn)
    R, W = an.shape[1], bn.shape[1]
    rm = torch.arange(R)[None, :] >= torch.tensor([l - 1 for l in il])[:, None]
    wm = torch.arange(W)[None, :] >= torch.tensor([l - 3 for l in cl])[:, None]
    A = A.masked_fill(rm[:, None, :, None] | wm[None, :, None, :], 0.0)
    S = {"MrSw": A.max(2)[0].sum(2), "MwSr": A.max(3)[0].sum(2)}[agg]
    (S * torch.tensor(Gup, dtype=torch.float64)).sum().backward()
    return S.detach().numpy(), a.grad.numpy(), b.grad.numpy()


@pytest.mark.parametrize("shape", [(6, 9, 35, 53, 64), (5, 7, 71, 71, 96), (4, 6, 20, 90, 64), (4, 5, 90, 40, 32)])
@pytest.mark.parametrize("agg", ["MrSw", "MwSr"])
def test_tiled_pair_backward_vs_autograd_and_generic(shape, agg, monkeypatch):
    """The register-tiled pair kernel (all three tile shapes, fused F.normalize Jacobian, gradients written in a
    permuted [S,B,d] layout) against float64 autograd and against the generic warp-per-word kernel."""
    from aladin_b200 import loss as L, synth
    Bi, Bc, S_im, S_s, d = shape
    im, s, il, cl = synth.raw_batch(77 + S_im, Bi, Bc, S_im, S_s, d, related=0.7)
    im[0, 1] = 0.0                                        # a zero token: norm below eps, no projection term
    Gup = np.random.RandomState(3).standard_normal((Bi, Bc)).astype(np.float32)
    Gup[np.random.RandomState(4).rand(Bi, Bc) < 0.3] = 0.0
    S_ref, dim_ref, ds_ref = _autograd_reference(im, s, il, cl, agg, Gup)

    def run(permuted):
        if permuted:
            im_t, s_t = cu(im.transpose(1, 0, 2).copy(), True), cu(s.transpose(1, 0, 2).copy(), True)
            a, b = im_t.permute(1, 0, 2), s_t.permute(1, 0, 2)
        else:
            im_t, s_t = cu(im, True), cu(s, True)
            a, b = im_t, s_t
        crit = L.AlignmentContrastiveLoss(aggregation=agg)
        crit.precision = "fp32"
        S = crit(a, b, il, cl, return_loss=False, return_similarity_mat=True)
        (S * cu(Gup)).sum().backward()
        gi, gs = im_t.grad.cpu().numpy(), s_t.grad.cpu().numpy()
        return S.detach().cpu().numpy(), (gi.transpose(1, 0, 2) if permuted else gi), (gs.transpose(1, 0, 2) if permuted else gs)

    S_t, dim_t, ds_t = run(permuted=True)
    assert_scores_close(S_t, S_ref, 1e-4, f"{agg} {shape}")
    np.testing.assert_allclose(dim_t, dim_ref, rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(ds_t, ds_ref, rtol=2e-3, atol=2e-5)
    monkeypatch.setenv("ALAD_BWD_GENERIC", "1")
    _, dim_g, ds_g = run(permuted=False)
    np.testing.assert_allclose(dim_t, dim_g, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ds_t, ds_g, rtol=1e-3, atol=1e-5)
