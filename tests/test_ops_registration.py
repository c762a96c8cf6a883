"""CPU-only: the torch custom-op layer registers every op under torch.ops.alad_b200 with a fake
(meta) kernel, so shapes propagate without touching the CUDA library; there is no CPU kernel."""
import pytest
import torch


def test_ops_are_registered_with_fake_kernels():
    from aladin_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.alad_b200, name), name
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        im = torch.empty((6, 9, 32), device="cuda")
        s = torch.empty((4, 12, 32), device="cuda")
        S = torch.ops.alad_b200.alignment_scores(im, s, [9] * 6, [12] * 4, "bf16", "MrSw")
        assert S.shape == (6, 4) and S.dtype == torch.float32
        d_im, d_s = torch.ops.alad_b200.alignment_scores_bwd(im, s, [9] * 6, [12] * 4, "MrSw", S)
        assert d_im.shape == im.shape and d_s.shape == s.shape
        M = torch.ops.alad_b200.dot_scores(im[:, 0], s[:, 0], "fp32")
        assert M.shape == (6, 4)
        sq = torch.empty((5, 5), device="cuda")
        loss, G = torch.ops.alad_b200.triplet(sq, 0.2, True)
        assert loss.shape == () and G.shape == (5, 5)
        loss, dM = torch.ops.alad_b200.listnet(sq, sq)
        assert loss.shape == () and dM.shape == (5, 5)
        r, t1 = torch.ops.alad_b200.rank_i2t(torch.empty((3, 15), device="cuda"), 5, 0)
        assert r.shape == (3,) and r.dtype == torch.int32 and t1.shape == (3,)
        r, tk = torch.ops.alad_b200.rank_t2i(torch.empty((60, 300), device="cuda"), 50, 5)
        assert r.shape == (300,) and tk.shape == (300, 50)


def test_no_cpu_kernel():
    from aladin_b200 import ops  # noqa: F401
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.alad_b200.triplet(torch.zeros(3, 3), 0.2, True)
