"""GPU: the torch.ops.alad_b200.* custom ops run the same kernels as the nn.Module drop-ins
(identical values) and their registered autograd matches the golden gradients."""
import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close, load_golden

pytestmark = pytest.mark.gpu


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


def test_alignment_op_forward_backward_golden():
    from aladin_b200 import ops  # noqa: F401
    g = load_golden("alignment_loss")
    im = cu(np.transpose(g["im_sbd"], (1, 0, 2)).copy(), True)
    s = cu(np.transpose(g["s_sbd"], (1, 0, 2)).copy(), True)
    S = torch.ops.alad_b200.alignment_scores(im, s, g["im_len"].tolist(), g["s_len"].tolist(), "fp32", "MrSw")
    assert_scores_close(S.detach().cpu().numpy(), g["S_mv"], 1e-4, "op scores")
    (S * cu(g["Gup"])).sum().backward()
    np.testing.assert_allclose(im.grad.cpu().numpy(), np.transpose(g["dim_dense"], (1, 0, 2)), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), np.transpose(g["ds_dense"], (1, 0, 2)), rtol=1e-3, atol=1e-5)


def test_loss_ops_match_modules():
    from aladin_b200 import loss as L, ops  # noqa: F401
    g = load_golden("triplet_listnet")
    S = cu(g["S"], True)
    loss, G = torch.ops.alad_b200.triplet(S, 0.2, True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["loss_mv"], rtol=1e-6)
    np.testing.assert_array_equal(S.grad.cpu().numpy(), g["G_mv"])
    T, M = cu(g["T"]), cu(g["M"], True)
    loss, _ = torch.ops.alad_b200.listnet(T, M)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["listnet_loss"], rtol=1e-5)
    np.testing.assert_allclose(M.grad.cpu().numpy(), g["listnet_dM"], rtol=2e-4, atol=1e-7)


def test_dot_and_rank_ops():
    from aladin_b200 import ops, ranking, scoring  # noqa: F401
    g = load_golden("matching")
    im, s = cu(g["im"], True), cu(g["s"], True)
    S = torch.ops.alad_b200.dot_scores(im, s, "fp32")
    assert_scores_close(S.detach().cpu().numpy(), g["S_dot_mv"], 1e-4, "dot op", floor=CANCELLING_FLOOR)
    loss, _ = torch.ops.alad_b200.triplet(S, 0.2, True)
    loss.backward()
    np.testing.assert_allclose(im.grad.cpu().numpy(), g["dim_dot_mv"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["ds_dot_mv"], rtol=1e-4, atol=1e-5)
    r = load_golden("retrieval")
    Sg = cu(r["S_full"])
    rank, top1 = torch.ops.alad_b200.rank_i2t(Sg, 5, 0)
    np.testing.assert_array_equal(rank.cpu().numpy(), r["ranks_i2t"])
    np.testing.assert_array_equal(top1.cpu().numpy(), r["top1"])
    rk, tk = torch.ops.alad_b200.rank_t2i(Sg, 50, 5)
    np.testing.assert_array_equal(rk.cpu().numpy(), r["ranks_t2i"])
