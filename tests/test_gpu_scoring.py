"""GPU parity tests for the packing + tcgen05 scoring kernels, through the C ABI.

Tolerances (BASELINE.json north_star): score matrices within 1e-2 absolute in
bf16-input / fp32-accumulate mode and within 1e-4 relative in fp32 mode."""
import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close, load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu

BF16_ATOL = 1e-2
FP32_RTOL = 1e-4


def test_pack_tokens_matches_normalize():
    from aladin_b200 import scoring, synth
    im, s, im_len, s_len = synth.raw_batch(1, 9, 7, 12, 15, 200)
    s[3, 2] = 0.0                                   # zero token -> stays zero (eps path)
    x = torch.from_numpy(s).cuda()
    nw = np.array([max(min(l - 3, 12), 0) for l in s_len], np.int32)
    for mode in (0, 1, 2):
        P = scoring.pack_tokens(x, nw, slot0=1, mode=mode, want_row_item=True)
        torch.cuda.synchronize()
        data = P.data.float().cpu().numpy()
        ri = P.row_item.cpu().numpy()
        xhat = O.l2_normalize(s)
        row = 0
        for b in range(7):
            for t in range(nw[b]):
                ref = xhat[b, 1 + t]
                got = data[row]
                assert ri[row] == b
                if mode == 0:
                    np.testing.assert_allclose(got[:200], ref, rtol=2 ** -8, atol=1e-30)
                    assert np.all(got[200:] == 0)
                else:
                    hi, second, third = got[:200], got[200:400], got[400:600]
                    lo, hi2 = (third, second) if mode == 1 else (second, third)
                    np.testing.assert_array_equal(hi, hi2)
                    np.testing.assert_allclose(hi + lo, ref, rtol=2 ** -15, atol=1e-30)
                    assert np.all(got[600:] == 0)
                row += 1
        assert row == P.n_rows and np.all(ri[row:] == -1)


@pytest.mark.parametrize("Ni,Nc,d", [(500, 300, 192), (240, 128, 64), (7, 1000, 1024), (1, 1, 64)])
def test_gemm_epilogue_is_bit_exact_on_exact_inputs(Ni, Nc, d):
    """Raw tcgen05 GEMM (TMA swizzle, smem/instruction descriptors, TMEM readback): inputs are
    small dyadic rationals so every product and partial sum is exact in fp32 -> bit-exact."""
    from aladin_b200 import scoring
    r = np.random.RandomState(7)
    im = (r.randint(-4, 5, size=(Ni, d)) / 8.0).astype(np.float32)
    s = (r.randint(-4, 5, size=(Nc, d)) / 8.0).astype(np.float32)
    got = scoring.dot_scores(torch.from_numpy(im).cuda(), torch.from_numpy(s).cuda(), precision="bf16")
    torch.cuda.synchronize()
    ref = im.astype(np.float64) @ s.astype(np.float64).T
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.float64), ref)


def test_mrsw_packed_is_bit_exact_on_exact_inputs():
    """Fused MrSw epilogue on un-normalised dyadic inputs: max and the ordered sums are exact."""
    from aladin_b200 import scoring, tiling
    r = np.random.RandomState(8)
    Bi, Bc, S_im, S_s, d = 41, 67, 35, 53, 128
    im = (r.randint(-4, 5, size=(Bi, S_im, d)) / 8.0).astype(np.float32)
    s = (r.randint(-4, 5, size=(Bc, S_s, d)) / 8.0).astype(np.float32)
    im_len = r.randint(1, S_im + 1, size=Bi); im_len[5] = S_im
    s_len = r.randint(3, S_s + 1, size=Bc); s_len[9] = S_s
    R, W, nr, nw, clamp = scoring.scored_counts(im.shape, s.shape, im_len, s_len)
    words = scoring.pack_tokens(torch.from_numpy(s).cuda(), nw, slot0=1, mode=0, normalize=False, want_row_item=True)
    regions = scoring.pack_tokens(torch.from_numpy(im).cuda(), nr, slot0=1, mode=0, normalize=False)
    _, table, _ = tiling.build_region_tiles(nr, clamp)
    tiles = torch.from_numpy(table.view(np.int32).reshape(-1).copy()).cuda()
    got = scoring.mrsw_scores_packed(words, regions, tiles, len(table), Bi, Bc)
    torch.cuda.synchronize()
    A = np.einsum("ird,jwd->ijrw", im[:, 1:].astype(np.float64), s[:, 1:-2].astype(np.float64))
    rm = np.arange(R)[None, :] >= nr[:, None]
    wm = np.arange(W)[None, :] >= nw[:, None]
    A[np.broadcast_to(rm[:, None, :, None] | wm[None, :, None, :], A.shape)] = 0
    ref = A.max(axis=2).sum(axis=2)
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.float64), ref)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_alignment_scores_golden(precision):
    import aladin_b200
    g = load_golden("alignment_scores")
    S = aladin_b200.alignment_scores(torch.from_numpy(g["im"]), torch.from_numpy(g["s"]), g["im_len"].tolist(),
                                     g["s_len"].tolist(), precision=precision)
    got = S.cpu().numpy()
    ref = g["S_MrSw"]
    if precision == "bf16":
        assert np.abs(got - ref).max() <= BF16_ATOL
    else:
        assert_scores_close(got, ref, FP32_RTOL)
    assert np.all(got[2] == 0) and np.all(got[:, 1] == 0)        # empty image / empty caption


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("shape", [(37, 53, 35, 53, 128), (64, 320, 71, 71, 1024), (130, 90, 35, 53, 768)])
def test_alignment_scores_vs_oracle(precision, shape):
    import aladin_b200
    from aladin_b200 import synth
    Bi, Bc, S_im, S_s, d = shape
    im, s, im_len, s_len = synth.raw_batch(3, Bi, Bc, S_im, S_s, d, related=0.5)
    # permuted [S,B,d] storage like alad_model.py:377-378
    im_t = torch.from_numpy(np.ascontiguousarray(im.transpose(1, 0, 2))).cuda().permute(1, 0, 2)
    s_t = torch.from_numpy(np.ascontiguousarray(s.transpose(1, 0, 2))).cuda().permute(1, 0, 2)
    got = aladin_b200.alignment_scores(im_t, s_t, im_len, s_len, precision=precision).cpu().numpy()
    ref = O.mrsw_scores(im, s, im_len, s_len, acc64=True)
    if precision == "bf16":
        assert np.abs(got - ref).max() <= BF16_ATOL, np.abs(got - ref).max()
    else:
        assert_scores_close(got, ref, FP32_RTOL)


def test_alignment_scores_dense_uniform_full_width():
    """BASELINE shape per pair (34 regions x 50 words, d=1024), all items full length: no clamp."""
    import aladin_b200
    from aladin_b200 import synth
    im, s, im_len, s_len = synth.raw_batch(4, 23, 31, 35, 53, 1024, ragged=False)
    got = aladin_b200.alignment_scores(torch.from_numpy(im), torch.from_numpy(s), im_len, s_len, precision="bf16")
    ref = O.mrsw_scores(im, s, im_len, s_len, acc64=True)
    assert np.abs(got.cpu().numpy() - ref).max() <= BF16_ATOL


def test_alignment_scores_is_deterministic():
    import aladin_b200
    from aladin_b200 import synth
    im, s, im_len, s_len = synth.raw_batch(5, 150, 400, 35, 53, 256)
    a = aladin_b200.alignment_scores(torch.from_numpy(im), torch.from_numpy(s), im_len, s_len)
    b = aladin_b200.alignment_scores(torch.from_numpy(im), torch.from_numpy(s), im_len, s_len)
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_all_pooling_modes_golden(precision):
    """alad/loss.py:120-135: sum, mean, MrSw, MrAVGw, symm, MwSr against the reference's outputs."""
    import aladin_b200
    g = load_golden("alignment_scores")
    for agg in ("sum", "mean", "MrSw", "MrAVGw", "symm", "MwSr"):
        S = aladin_b200.alignment_scores(torch.from_numpy(g["im"]), torch.from_numpy(g["s"]), g["im_len"].tolist(),
                                         g["s_len"].tolist(), precision=precision, aggregation=agg).cpu().numpy()
        ref = g["S_" + agg]
        ok = np.isfinite(ref)                       # MrAVGw divides by a zero word count for one caption
        assert np.array_equal(np.isnan(S), np.isnan(ref)), agg
        if precision == "bf16" and agg not in ("sum", "mean"):
            assert np.abs(S[ok] - ref[ok]).max() <= 2 * BF16_ATOL, (agg, np.abs(S[ok] - ref[ok]).max())
        else:
            assert_scores_close(S[ok], ref[ok], FP32_RTOL, agg)


@pytest.mark.parametrize("agg", ["MwSr", "symm", "MrAVGw", "sum", "mean"])
def test_pooling_modes_vs_oracle_ragged(agg):
    import aladin_b200
    from aladin_b200 import synth
    im, s, im_len, s_len = synth.raw_batch(11, 45, 38, 30, 41, 192, related=0.5)
    got = aladin_b200.alignment_scores(torch.from_numpy(im), torch.from_numpy(s), im_len, s_len, precision="fp32",
                                       aggregation=agg).cpu().numpy()
    ref = O.alignment_scores_small(im, s, im_len, s_len, agg)
    ok = np.isfinite(ref)
    assert_scores_close(got[ok], ref[ok], FP32_RTOL, agg, floor=CANCELLING_FLOOR if agg in ("sum", "mean") else 0.01)


@pytest.mark.parametrize("shape", [(37, 53, 35, 53, 128), (64, 320, 71, 71, 1024), (130, 90, 35, 53, 768)])
def test_tf32_mode_gallery_scores_vs_oracle(shape):
    """The 'tf32' precision of the retrieval gallery (fp32 operands rounded to TF32, tcgen05 kind::tf32): an intermediate
    mode -- half the cost of the split-precision 'fp32' mode, ten times tighter than 'bf16'.  Its worst entry sits at
    2.6e-4 .. 5.6e-4 relative (1 % floor) on these shapes (tools/tf32_diag.py), so it does NOT meet the 1e-4 of the parity
    mode ('fp32': <= 1e-5 here); the bound asserted for it is 1e-3.  Ragged lengths, an image without regions."""
    from aladin_b200 import retrieval, synth
    Bi, Bc, S_im, S_s, d = shape
    im, s, il, cl = synth.raw_batch(21, Bi, Bc, S_im, S_s, d, related=0.5)
    il[3] = 1
    gal = retrieval.AlignmentGallery(torch.from_numpy(im).cuda(), torch.from_numpy(s).cuda(), il, cl, n_images=Bi, precision="tf32")
    got = gal.scores().cpu().numpy()
    ref = O.mrsw_scores(im, s, il, cl, acc64=True)
    assert_scores_close(got, ref, 1e-3, "tf32 gallery scores")
    # and it is a different (cheaper) path than the split-precision one: not bit-identical, but as close
    g32 = retrieval.AlignmentGallery(torch.from_numpy(im).cuda(), torch.from_numpy(s).cuda(), il, cl, n_images=Bi, precision="fp32")
    assert_scores_close(g32.scores().cpu().numpy(), ref, FP32_RTOL, "fp32 gallery scores")
