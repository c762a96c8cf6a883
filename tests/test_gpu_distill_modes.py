"""GPU parity tests for the remaining DistillationLoss modes (alad/loss.py:371-425), order_sim,
the cosine_sim gradient and the 'sum' / 'mean' pooling gradients: golden vectors produced by the
unmodified reference + the oracle at larger sizes.  Integer-valued gradients are compared exactly."""
import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close, load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


def test_mse_golden():
    from aladin_b200 import loss as L
    g = load_golden("distill_modes")
    dl = L.DistillationLoss(mode="mse").cuda()
    assert list(dl.state_dict().keys()) == ["wb"]                    # checkpoint key kept (alad/loss.py:367)
    with torch.no_grad():
        dl.wb.copy_(torch.tensor(g["mse_wb"]))
    T, M = cu(g["T"], True), cu(g["M"], True)
    loss = dl(T, M)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["mse_loss"], rtol=1e-5)
    np.testing.assert_allclose(M.grad.cpu().numpy(), g["mse_dM"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(dl.wb.grad.cpu().numpy(), g["mse_dwb"], rtol=1e-4, atol=1e-7)
    assert T.grad is None                                            # teacher detached (alad/loss.py:370)


@pytest.mark.parametrize("margin", [0.2, 0.05])
def test_contrastive_golden(margin):
    from aladin_b200 import loss as L
    g = load_golden("distill_modes")
    T, M = cu(g["T"]), cu(g["M"], True)
    loss = L.DistillationLoss(mode="contrastive", margin=margin)(T, M)
    loss.backward()
    k = f"contrastive_m{margin}"
    np.testing.assert_allclose(loss.item(), g[k + "_loss"], rtol=1e-5)
    np.testing.assert_array_equal(M.grad.cpu().numpy(), g[k + "_dM"])
    np.testing.assert_array_equal(T.cpu().numpy(), g[k + "_T_after"])  # in-place diagonal zeroing, like the reference


@pytest.mark.parametrize("margin,thr,stride", [(0.2, 0.1, 3), (0.1, 0.5, 1)])
def test_ordinal_golden(margin, thr, stride):
    from aladin_b200 import loss as L
    g = load_golden("distill_modes")
    T, M = cu(g["T"]), cu(g["M"], True)
    loss = L.DistillationLoss(mode="ordinal", margin=margin, threshold=thr, stride=stride)(T, M)
    loss.backward()
    k = f"ordinal_m{margin}_t{thr}_s{stride}"
    np.testing.assert_allclose(loss.item(), g[k + "_loss"], rtol=1e-5)
    np.testing.assert_allclose(M.grad.cpu().numpy(), g[k + "_dM"], rtol=1e-5, atol=1e-9)


def test_ordinal_empty_selection_is_nan():
    from aladin_b200 import loss as L
    g = load_golden("distill_modes")
    loss, dM = L.distill_ordinal_fwd_bwd(cu(g["T"]), cu(g["M"]), 0.2, 100.0, 3)
    assert np.isnan(loss.item()) and np.isnan(g["ordinal_m0.2_t100.0_s3_loss"])
    assert float(dM.abs().max()) == 0.0                              # autograd of an empty mean: no gradient


@pytest.mark.parametrize("B", [2, 33, 512, 1100])
def test_distill_modes_vs_oracle(B):
    from aladin_b200 import loss as L
    r = np.random.RandomState(100 + B)
    T = (r.standard_normal((B, B)) * 0.7 + 0.3).astype(np.float32)
    M = np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32)
    wb = np.array([0.9, -0.1], np.float32)
    loss, dM, dwb = L.distill_mse_fwd_bwd(cu(T), cu(M), cu(wb))
    rl, rM, rwb = O.distill_mse(T, M, wb)
    np.testing.assert_allclose(loss.item(), rl, rtol=2e-5)
    np.testing.assert_allclose(dM.cpu().numpy(), rM, rtol=1e-4, atol=1e-7 * np.abs(rM).max())
    np.testing.assert_allclose(dwb.cpu().numpy(), rwb, rtol=1e-3, atol=1e-5)
    loss, dM = L.distill_contrastive_fwd_bwd(cu(T), cu(M), 0.2)
    rl, rM = O.distill_contrastive(T, M, 0.2)
    np.testing.assert_allclose(loss.item(), rl, rtol=2e-5)
    np.testing.assert_array_equal(dM.cpu().numpy(), rM)
    for stride in (1, 3):
        if stride >= B:
            continue
        loss, dM = L.distill_ordinal_fwd_bwd(cu(T), cu(M), 0.2, 0.1, stride)
        rl, rM = O.distill_ordinal(T, M, 0.2, 0.1, stride)
        np.testing.assert_allclose(loss.item(), rl, rtol=2e-5)
        np.testing.assert_allclose(dM.cpu().numpy(), rM, rtol=1e-5, atol=1e-9)


def test_ordinal_ties_are_stable():
    """Teacher rows with repeated values: the CUDA sort is stable (index ascending), like the oracle."""
    from aladin_b200 import loss as L
    r = np.random.RandomState(3)
    B = 40
    T = r.randint(0, 4, size=(B, B)).astype(np.float32)
    M = r.standard_normal((B, B)).astype(np.float32)
    loss, dM = L.distill_ordinal_fwd_bwd(cu(T), cu(M), 0.2, 1.0, 2)
    rl, rM = O.distill_ordinal(T, M, 0.2, 1.0, 2)
    np.testing.assert_allclose(loss.item(), rl, rtol=1e-5)
    np.testing.assert_allclose(dM.cpu().numpy(), rM, rtol=1e-5, atol=1e-9)


def test_order_sim_golden_and_gradient():
    from aladin_b200 import loss as L
    g = load_golden("distill_modes")
    S = L.order_sim(cu(g["order_im"]), cu(g["order_s"]))
    np.testing.assert_allclose(S.cpu().numpy(), g["order_S"], rtol=1e-5, atol=1e-6)
    # gradient against torch autograd on the reference formula (alad/loss.py:20-26)
    r = np.random.RandomState(4)
    im, s = r.standard_normal((70, 100)).astype(np.float32), r.standard_normal((90, 100)).astype(np.float32)
    Gup = r.standard_normal((70, 90)).astype(np.float32)
    a, b = cu(im, True), cu(s, True)
    out = L.order_sim(a, b)
    (out * cu(Gup)).sum().backward()
    at = torch.tensor(im, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(s, dtype=torch.float64, requires_grad=True)
    ref = -(bt.unsqueeze(1) - at.unsqueeze(0)).clamp(min=0).pow(2).sum(2).sqrt().t()
    (ref * torch.tensor(Gup, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(a.grad.cpu().numpy(), at.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(b.grad.cpu().numpy(), bt.grad.numpy(), rtol=1e-4, atol=1e-5)
    # measure='order' is constructible like in the reference (alad/loss.py:36-37)
    loss = L.ContrastiveLoss(margin=0.2, measure="order", max_violation=True)(cu(g["order_im"][:7]), cu(g["order_s"][:7]))
    assert np.isfinite(loss.item())


@pytest.mark.parametrize("key,mv", [("cosine_mv", True), ("cosine_sum", False)])
def test_cosine_measure_gradient_golden(key, mv):
    import aladin_b200
    from aladin_b200 import loss as L
    g = load_golden("matching")
    aladin_b200.set_precision("fp32")
    try:
        im, s = cu(g["im"] * 2.5, True), cu(g["s"], True)
        loss, S = L.ContrastiveLoss(margin=0.2, measure="cosine", max_violation=mv)(im, s, return_similarity_mat=True)
        loss.backward()
        assert_scores_close(S.detach().cpu().numpy(), g["S_" + key], 1e-4, key, floor=CANCELLING_FLOOR)
        np.testing.assert_allclose(loss.item(), g["loss_" + key], rtol=1e-4)
        np.testing.assert_allclose(im.grad.cpu().numpy(), g["dim_" + key], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(s.grad.cpu().numpy(), g["ds_" + key], rtol=1e-3, atol=1e-5)
    finally:
        aladin_b200.set_precision("bf16")


@pytest.mark.parametrize("agg", ["sum", "mean", "MrAVGw", "MwSr", "symm"])
def test_pooling_mode_gradients_golden(agg):
    """Forward + gradients of every non-default pooling mode against the reference's autograd."""
    from aladin_b200 import loss as L
    g = load_golden("pooled_grads")
    im, s = cu(g["im"], True), cu(g["s"], True)
    crit = L.AlignmentContrastiveLoss(aggregation=agg)
    crit.precision = "fp32"
    S = crit(im, s, g["im_len"].tolist(), g["s_len"].tolist(), return_loss=False, return_similarity_mat=True)
    (S * cu(g["Gup"])).sum().backward()
    assert_scores_close(S.detach().cpu().numpy(), g["S_" + agg], 1e-4, agg, floor=CANCELLING_FLOOR if agg in ("sum", "mean") else 0.01)
    scale = np.abs(g["dim_" + agg]).max()
    np.testing.assert_allclose(im.grad.cpu().numpy(), g["dim_" + agg], rtol=2e-3, atol=2e-5 * max(scale, 1e-3))
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["ds_" + agg], rtol=2e-3, atol=2e-5 * max(scale, 1e-3))
