"""CPU-only: the CUDA-core translation units of aladin_b200/csrc (losses, distill, rank, misc_sim, pack, mrsw_bwd --
every kernel of the path that is not tcgen05 / TMA) compiled for the host-thread emulator of tests/cuda_emu
and called through their C-ABI entry points on numpy buffers, against the oracle.  Same source as the GPU build;
small sizes (one std::thread per CUDA thread).  The `-m gpu` tests remain the parity tests proper; this file
catches index / barrier mistakes on the GPU-less box, and the ThreadSanitizer test at the end reports any
shared-memory hand-off that lacks a barrier."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import alad_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuda_emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
TSAN = os.environ.get("ALAD_EMU_TSAN") == "1"      # inside the sanitizer child run: same kernels, smallest shapes only


def big(*cases):
    """Parameter sets that only run outside the (10 x slower) sanitizer child."""
    return [] if TSAN else list(cases)


def load_emu(name, tsan=False):
    """ctypes handle of the emulated translation unit with the prototypes of aladin_b200/_cabi.py."""
    import build_emu
    from aladin_b200 import _cabi
    # one emulated library with every translation unit (built once, in parallel); `name` documents the unit under test
    lib = C.CDLL(build_emu.build_library(tsan=tsan or os.environ.get("ALAD_EMU_TSAN") == "1"))
    for sym, (res, args) in _cabi.PROTOTYPES.items():
        try:
            fn = getattr(lib, sym)
        except AttributeError:
            continue
        fn.restype, fn.argtypes = res, args
    return lib


@pytest.fixture(scope="module")
def losses():
    return load_emu("losses")


@pytest.fixture(scope="module")
def distill():
    return load_emu("distill")


@pytest.fixture(scope="module")
def rank():
    return load_emu("rank")


@pytest.fixture(scope="module")
def misc():
    return load_emu("misc_sim")


@pytest.fixture(scope="module")
def pack():
    return load_emu("pack")


def p(a):
    return a.ctypes.data if a is not None else None


def ok(lib, rc):
    assert rc == 0, lib.alad_last_error()


def workspace(nbytes):
    return np.zeros(max(int(nbytes), 16) // 4 + 4, np.float32)


# ------------------------------------------------------------------------------------------------- losses.cu
def run_triplet(lib, S, margin, mv):
    B = S.shape[0]
    loss = np.zeros(1, np.float32)
    G = np.full((B, B), np.nan, np.float32)
    ra, ca = np.zeros(B, np.int32), np.zeros(B, np.int32)
    ws = workspace(lib.alad_loss_workspace_bytes(B))
    ok(lib, lib.alad_triplet_fwd_bwd(p(S), B, B, margin, 1 if mv else 0, p(loss), p(G), B, p(ra), p(ca), p(ws), None))
    return loss[0], G


@pytest.mark.parametrize("B", [1, 7] + big(45))
@pytest.mark.parametrize("mv", [True, False])
def test_triplet_emulated(losses, B, mv):
    r = np.random.RandomState(B)
    S = r.standard_normal((B, B)).astype(np.float32)
    loss, G = run_triplet(losses, S, 0.2, mv)
    np.testing.assert_allclose(loss, O.triplet_loss(S, 0.2, mv), rtol=2e-5)
    np.testing.assert_array_equal(G, O.triplet_grad(S, 0.2, mv))


def test_triplet_emulated_golden(losses):
    from conftest import load_golden
    g = load_golden("triplet_listnet")
    for key, mv in (("mv", True), ("sum", False)):
        loss, G = run_triplet(losses, np.ascontiguousarray(g["S"]), 0.2, mv)
        np.testing.assert_allclose(loss, g[f"loss_{key}"], rtol=1e-6)
        np.testing.assert_array_equal(G, g[f"G_{key}"])


@pytest.mark.parametrize("B", [1, 9] + big(40, 136, 150))      # 136 / 150: two column strips x five row chunks
def test_listnet_emulated(losses, B):
    r = np.random.RandomState(B)
    T = (r.standard_normal((B, B)) * 2 + 3).astype(np.float32)
    M = np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32)
    loss = np.zeros(1, np.float32)
    dM = np.full((B, B), np.nan, np.float32)
    ws = workspace(losses.alad_loss_workspace_bytes(B))
    ok(losses, losses.alad_listnet_fwd_bwd(p(T), B, p(M), B, B, 6.0, 1e-10, p(loss), p(dM), B, p(ws), None))
    np.testing.assert_allclose(loss[0], O.listnet_loss(T, M), rtol=2e-5)
    ref = O.listnet_grad(T, M)
    np.testing.assert_allclose(dM, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())


# ------------------------------------------------------------------------------------------------- distill.cu
def _distill_inputs(B, seed):
    r = np.random.RandomState(seed)
    T = (r.standard_normal((B, B)) * 2 + 3).astype(np.float32)
    M = np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32)
    return T, M


@pytest.mark.parametrize("B", [4] + big(37))
def test_distill_mse_emulated(distill, B):
    T, M = _distill_inputs(B, B)
    wb = np.array([0.7, -0.2], np.float32)
    loss, dM, dwb = np.zeros(1, np.float32), np.zeros((B, B), np.float32), np.zeros(2, np.float32)
    ws = workspace(distill.alad_distill_workspace_bytes(B, 0))
    ok(distill, distill.alad_distill_mse_fwd_bwd(p(T), B, p(M), B, B, p(wb), p(loss), p(dM), B, p(dwb), p(ws), None))
    rl, rdM, rdwb = O.distill_mse(T, M, wb)
    np.testing.assert_allclose(loss[0], rl, rtol=2e-5)
    np.testing.assert_allclose(dM, rdM, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dwb, rdwb, rtol=1e-4)


@pytest.mark.parametrize("B", [5] + big(29))
def test_distill_contrastive_emulated(distill, B):
    T, M = _distill_inputs(B, 100 + B)
    Tc = T.copy()
    loss, dM = np.zeros(1, np.float32), np.zeros((B, B), np.float32)
    ws = workspace(distill.alad_distill_workspace_bytes(B, 1))
    ok(distill, distill.alad_distill_contrastive_fwd_bwd(p(Tc), B, p(M), B, B, 0.2, 1, p(loss), p(dM), B, p(ws), None))
    rl, rG = O.distill_contrastive(T, M, 0.2)
    np.testing.assert_allclose(loss[0], rl, rtol=2e-5)
    np.testing.assert_allclose(dM, rG, rtol=1e-5, atol=1e-7)
    assert np.all(np.diag(Tc) == 0) and np.array_equal(Tc - np.diag(np.diag(Tc)), T - np.diag(np.diag(T)))


@pytest.mark.parametrize("B,stride", [(6, 1)] + big((29, 3)))
def test_distill_ordinal_emulated(distill, B, stride):
    T, M = _distill_inputs(B, 200 + B)
    loss, dM = np.zeros(1, np.float32), np.zeros((B, B), np.float32)
    ws = workspace(distill.alad_distill_workspace_bytes(B, 2))
    ok(distill, distill.alad_distill_ordinal_fwd_bwd(p(T), B, p(M), B, B, 0.2, 0.1, stride, p(loss), p(dM), B, p(ws), None))
    rl, rG = O.distill_ordinal(T, M, 0.2, 0.1, stride)
    np.testing.assert_allclose(loss[0], rl, rtol=2e-5)
    np.testing.assert_allclose(dM, rG, rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------------- rank.cu
def _stable_desc(v):
    return np.argsort(v, kind="stable")[::-1]


@pytest.mark.parametrize("Ni,ties,k", [(12, False, 10), (14, True, 10)] + big((40, True, 30)))
def test_ranking_emulated(rank, Ni, ties, k):
    r = np.random.RandomState(Ni)
    Nc = 5 * Ni
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    S[np.arange(Nc) // 5, np.arange(Nc)] += 1.0
    if ties:
        S = np.round(S * 2) / 2
    rk, top1 = np.zeros(Ni, np.int32), np.zeros(Ni, np.int32)
    ok(rank, rank.alad_rank_rows(p(S), Nc, Ni, Nc, 5, 0, p(rk), p(top1), None))
    gt = np.zeros(Nc, np.float32)
    ok(rank, rank.alad_col_gt(p(S), Nc, Ni, Nc, 5, 0, p(gt), None))
    cnt = np.zeros(Nc, np.int32)
    ok(rank, rank.alad_col_count(p(S), Nc, Ni, Nc, 5, 0, p(gt), p(cnt), None))
    sel_s, sel_i = np.zeros((Nc, k), np.float32), np.zeros((Nc, k), np.int32)
    ws = workspace(rank.alad_col_topk_select_workspace_bytes(Ni, Nc, k))
    ok(rank, rank.alad_col_topk_select(p(S), Nc, Ni, Nc, k, 0, p(sel_s), p(sel_i), p(ws), None))
    # heap variant in two row slices + merge
    splits = 2
    cs, ci = np.zeros((splits, Nc, k), np.float32), np.zeros((splits, Nc, k), np.int32)
    ok(rank, rank.alad_col_topk(p(S), Nc, Ni, Nc, k, 0, splits, p(cs), p(ci), None))
    ms, mi = np.zeros((Nc, k), np.float32), np.zeros((Nc, k), np.int32)
    ok(rank, rank.alad_topk_merge(p(cs), p(ci), splits, Nc, k, p(ms), p(mi), None))

    kk = min(k, Ni)
    for i in range(Ni):
        inds = _stable_desc(S[i])
        pos = np.empty(Nc, np.int64)
        pos[inds] = np.arange(Nc)
        assert rk[i] == pos[5 * i:5 * i + 5].min() and top1[i] == inds[0]
    for c in range(Nc):
        inds = _stable_desc(S[:, c])
        assert gt[c] == S[c // 5, c]
        assert cnt[c] == np.where(inds == c // 5)[0][0]
        np.testing.assert_array_equal(sel_i[c, :kk], inds[:kk])
        np.testing.assert_array_equal(mi[c, :kk], inds[:kk])
        np.testing.assert_array_equal(sel_s[c, :kk], S[inds[:kk], c])
        assert np.all(sel_i[c, kk:] == -1) and np.all(mi[c, kk:] == -1)
    if not ties:
        ri, t1 = O.i2t_ranks(S)
        np.testing.assert_array_equal(rk, ri)
        np.testing.assert_array_equal(top1, t1)


def _fused(rank, S, Nc, q_rows, q_cols, k, img_off, gt, want_count):
    Ni = S.shape[0]
    ld = S.strides[0] // 4
    rk, top1 = np.full(q_rows, -7, np.int32), np.full(q_rows, -7, np.int32)
    cnt = np.full(q_cols, -7, np.int32) if want_count else None
    ts, ti = np.zeros((q_cols, k), np.float32), np.zeros((q_cols, k), np.int32)
    ws = workspace(rank.alad_rank_fused_workspace_bytes(Ni, q_rows, q_cols, k))
    ok(rank, rank.alad_rank_fused(p(S), ld, Ni, Nc, 5, img_off, q_rows, q_cols, k, p(gt), p(rk), p(top1), p(cnt), p(ts), p(ti),
                                  p(ws), None))
    return rk, top1, cnt, ts, ti


@pytest.mark.parametrize("Ni,Nc,q_rows,q_cols,k,img_off,mode", [(264, 60, 264, 60, 10, 0, "own")] + big(
    (264, 2052, 200, 400, 10, 3, "given"), (260, 52, 260, 52, 10, 0, "nocount"),
    (260, 48, 260, 48, 10, 0, "masked"), (40, 61, 40, 61, 10, 0, "own"),
    (264, 60, 0, 60, 10, 0, "own")))                    # no i2t queries (ranking.t2i_rank_topk)
def test_rank_fused_emulated(rank, Ni, Nc, q_rows, q_cols, k, img_off, mode):
    """alad_rank_fused (one sweep for rows + counts + group maxima, one for the candidates) == the one-purpose entry
    points on the same block: query sub-ranges, an image offset, exact ties, -inf (masked) scores, a column count that is
    not a multiple of four (unaligned: falls back), blocks too small for the threshold select (fall back)."""
    r = np.random.RandomState(Ni + Nc)
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    if mode == "ties":
        S = np.round(S * 2) / 2
    if mode == "masked":
        S[r.rand(Ni, Nc) < 0.6] = -np.inf
        S[5] = -np.inf
    gt = None
    if mode in ("given", "nocount"):
        gt = r.standard_normal(q_cols).astype(np.float32)
        gt[::3] = S[(np.arange(q_cols) // 5) % Ni, np.arange(q_cols)][::3]
    rk, top1, cnt, ts, ti = _fused(rank, S, Nc, q_rows, q_cols, k, img_off, gt if mode == "given" else None, mode != "nocount")
    if TSAN:        # the sanitizer run is about the fused kernels' barriers: numpy is the checker, not the one-purpose kernels
        for i in range(q_rows):
            inds = _stable_desc(S[i])
            pos = np.empty(Nc, np.int64)
            pos[inds] = np.arange(Nc)
            assert top1[i] == inds[0] and rk[i] == (pos[5 * i:5 * i + 5].min() if 5 * i < Nc else Nc)
        for c in range(q_cols):
            inds = _stable_desc(S[:, c])
            assert cnt[c] == np.where(inds == c // 5)[0][0]
            np.testing.assert_array_equal(ti[c], inds[:k])
        return
    # the one-purpose kernels on the same block
    rk0, top10 = np.zeros(q_rows, np.int32), np.zeros(q_rows, np.int32)
    ok(rank, rank.alad_rank_rows(p(S), Nc, q_rows, Nc, 5, img_off, p(rk0), p(top10), None))
    np.testing.assert_array_equal(rk, rk0)
    np.testing.assert_array_equal(top1, top10)
    if mode != "nocount":
        if mode != "given":
            gt = np.zeros(q_cols, np.float32)
            ok(rank, rank.alad_col_gt(p(S), Nc, Ni, q_cols, 5, img_off, p(gt), None))
        cnt0 = np.zeros(q_cols, np.int32)
        ok(rank, rank.alad_col_count(p(S), Nc, Ni, q_cols, 5, img_off, p(gt), p(cnt0), None))
        np.testing.assert_array_equal(cnt, cnt0)
    s0, i0 = np.zeros((q_cols, k), np.float32), np.zeros((q_cols, k), np.int32)
    ws = workspace(rank.alad_col_topk_select_workspace_bytes(Ni, q_cols, k))
    ok(rank, rank.alad_col_topk_select(p(S), Nc, Ni, q_cols, k, img_off, p(s0), p(i0), p(ws), None))
    np.testing.assert_array_equal(ti, i0)
    np.testing.assert_array_equal(ts, s0)
    # and numpy's stable argsort reversed, the reference's order (alad/evaluation.py:213-223, 303-308)
    for i in range(0, q_rows, 37):
        inds = _stable_desc(S[i])
        assert top1[i] == inds[0]
        g0 = 5 * (img_off + i)
        if g0 < Nc:
            pos = np.empty(Nc, np.int64)
            pos[inds] = np.arange(Nc)
            assert rk[i] == pos[g0:g0 + 5].min()
        else:
            assert rk[i] == Nc
    for c in range(0, q_cols, 11):
        inds = _stable_desc(S[:, c])
        np.testing.assert_array_equal(ti[c], inds[:k] + img_off)


def test_ranking_emulated_reference_golden(rank):
    from conftest import load_golden
    g = load_golden("retrieval")
    S = np.ascontiguousarray(g["S_full"], np.float32)
    Ni, Nc = S.shape
    rk, top1 = np.zeros(Ni, np.int32), np.zeros(Ni, np.int32)
    ok(rank, rank.alad_rank_rows(p(S), Nc, Ni, Nc, 5, 0, p(rk), p(top1), None))
    np.testing.assert_array_equal(rk, g["ranks_i2t"])
    np.testing.assert_array_equal(top1, g["top1"])
    gt, cnt = np.zeros(Nc, np.float32), np.zeros(Nc, np.int32)
    ok(rank, rank.alad_col_gt(p(S), Nc, Ni, Nc, 5, 0, p(gt), None))
    ok(rank, rank.alad_col_count(p(S), Nc, Ni, Nc, 5, 0, p(gt), p(cnt), None))
    np.testing.assert_array_equal(cnt, g["ranks_t2i"])
    k = 50
    sel_s, sel_i = np.zeros((Nc, k), np.float32), np.zeros((Nc, k), np.int32)
    ws = workspace(rank.alad_col_topk_select_workspace_bytes(Ni, Nc, k))
    ok(rank, rank.alad_col_topk_select(p(S), Nc, Ni, Nc, k, 0, p(sel_s), p(sel_i), p(ws), None))
    np.testing.assert_array_equal(sel_i, g["top50"])


# ------------------------------------------------------------------------------------------------- misc_sim.cu
def test_order_sim_emulated(misc):
    r = np.random.RandomState(3)
    Ni, Nc, d = 70, 33, 40
    im = np.abs(r.standard_normal((Ni, d))).astype(np.float32)
    s = np.abs(r.standard_normal((Nc, d))).astype(np.float32)
    out = np.zeros((Ni, Nc), np.float32)
    ok(misc, misc.alad_order_scores(p(im), d, p(s), d, Ni, Nc, d, p(out), Nc, None))
    np.testing.assert_allclose(out, O.order_scores(im, s), rtol=2e-5, atol=1e-6)
    G = r.standard_normal((Ni, Nc)).astype(np.float32)
    d_im, d_s = np.zeros_like(im), np.zeros_like(s)
    ok(misc, misc.alad_order_scores_bwd(p(im), d, p(s), d, Ni, Nc, d, p(out), Nc, p(G), Nc, p(d_im), p(d_s), None))
    diff = np.maximum(s[None].astype(np.float64) - im[:, None], 0.0)                  # [Ni,Nc,d]
    w = np.where(out != 0, G / np.where(out != 0, out, 1), 0.0).astype(np.float64)
    np.testing.assert_allclose(d_s, np.einsum("ij,ijk->jk", w, diff), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(d_im, -np.einsum("ij,ijk->ik", w, diff), rtol=1e-4, atol=1e-6)


def _normalize_jacobian(x, g, eps):
    x, g = x.astype(np.float64), g.astype(np.float64)
    n = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), eps)
    xh = x / n
    return (g - xh * (xh * g).sum(-1, keepdims=True)) / n


def test_normalize_and_pool_backward_emulated(misc):
    r = np.random.RandomState(4)
    rows, d = 37, 100
    x = r.standard_normal((rows, d)).astype(np.float32)
    dx = r.standard_normal((rows, d)).astype(np.float32)
    ref = _normalize_jacobian(x, dx, 1e-12)
    ok(misc, misc.alad_normalize_bwd(p(x), d, rows, d, 1e-12, p(dx), d, None))
    np.testing.assert_allclose(dx, ref, rtol=1e-4, atol=1e-6)
    B, S, d = 6, 9, 48
    src = r.standard_normal((B, S, d)).astype(np.float32)
    cnt = np.array([8, 0, 3, 8, 1, 5], np.int32)
    d_pool = r.standard_normal((B, d)).astype(np.float32)
    out = np.full((B, S, d), np.nan, np.float32)
    ok(misc, misc.alad_pool_tokens_bwd(p(src), S * d, d, B, S, d, 1, p(cnt), 1e-12, p(d_pool), p(out), None))
    ref = np.zeros((B, S, d))
    for b in range(B):
        for t in range(cnt[b]):
            ref[b, 1 + t] = _normalize_jacobian(src[b, 1 + t][None], d_pool[b][None], 1e-12)[0]
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------- pack.cu
def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("d,mode", [(200, 0), (200, 1), (200, 2), (37, 0), (37, 1)])
def test_pack_tokens_emulated(pack, d, mode):
    from aladin_b200._cabi import PackArgs
    r = np.random.RandomState(d + mode)
    B, S = 5, 9
    x = r.standard_normal((B, S, d)).astype(np.float32)
    x[2, 3] = 0.0                                               # zero token stays zero (eps path)
    cnt = np.array([8, 0, 4, 1, 6], np.int32)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    rows = int(cnt.sum())
    Kp = -(-(d * (1 if mode == 0 else 3)) // 64) * 64
    dst = np.full((rows, Kp), 0x7FC0, np.uint16)
    item = np.full(rows, -7, np.int32)
    a = PackArgs(src=p(x), stride_b=S * d, stride_s=d, B=B, S=S, d=d, slot0=1, count=p(cnt), row_off=p(off), dst=p(dst),
                 Kp=Kp, mode=mode, normalize=1, eps=1e-12, row_item=p(item), item_base=100)
    ok(pack, pack.alad_pack_tokens(C.byref(a), None))
    got = _bf16_to_f32(dst)
    xh = O.l2_normalize(x)
    row = 0
    for b in range(B):
        for t in range(cnt[b]):
            ref = xh[b, 1 + t]
            assert item[row] == 100 + b
            if mode == 0:
                np.testing.assert_allclose(got[row, :d], ref, rtol=2 ** -8, atol=1e-30)
                assert np.all(got[row, d:] == 0)
            else:
                hi, second, third = got[row, :d], got[row, d:2 * d], got[row, 2 * d:3 * d]
                lo, hi2 = (third, second) if mode == 1 else (second, third)
                np.testing.assert_array_equal(hi, hi2)
                np.testing.assert_allclose(hi + lo, ref, rtol=2 ** -15, atol=1e-30)
                assert np.all(got[row, 3 * d:] == 0)
            row += 1
    assert row == rows


@pytest.mark.parametrize("B,d,mode", [(19, 96, 0), (8, 40, 2), (1, 64, 1)])
def test_pack_single_slot_items_emulated(pack, B, d, mode):
    """One slot per item (the plain-GEMM operands): one item per warp, B not a multiple of the 8 warps."""
    from aladin_b200._cabi import PackArgs
    r = np.random.RandomState(B + d)
    x = r.standard_normal((B, 1, d)).astype(np.float32)
    cnt = np.ones(B, np.int32)
    cnt[B // 2] = 0 if B > 1 else 1
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    rows = int(cnt.sum())
    Kp = -(-(d * (1 if mode == 0 else 3)) // 64) * 64
    dst = np.full((rows, Kp), 0x7FC0, np.uint16)
    a = PackArgs(src=p(x), stride_b=d, stride_s=d, B=B, S=1, d=d, slot0=0, count=p(cnt), row_off=p(off), dst=p(dst), Kp=Kp,
                 mode=mode, normalize=0, eps=0.0, row_item=None, item_base=0)
    ok(pack, pack.alad_pack_tokens(C.byref(a), None))
    got = _bf16_to_f32(dst)
    assert not np.isnan(got).any()
    row = 0
    for b in range(B):
        if not cnt[b]:
            continue
        hi = got[row, :d]
        np.testing.assert_allclose(hi, x[b, 0], rtol=2 ** -8, atol=1e-30)
        if mode:
            lo = got[row, 2 * d:3 * d] if mode == 1 else got[row, d:2 * d]
            np.testing.assert_allclose(hi + lo, x[b, 0], rtol=2 ** -15, atol=1e-30)
        row += 1
    assert row == rows


def test_pool_and_scale_emulated(pack):
    r = np.random.RandomState(9)
    B, S, d = 4, 7, 50
    x = r.standard_normal((B, S, d)).astype(np.float32)
    cnt = np.array([6, 2, 0, 4], np.int32)
    out = np.full((B, d), np.nan, np.float32)
    ok(pack, pack.alad_pool_tokens(p(x), S * d, d, B, S, d, 1, p(cnt), 1e-12, p(out), None))
    xh = O.l2_normalize(x).astype(np.float64)
    ref = np.stack([xh[b, 1:1 + cnt[b]].sum(0) for b in range(B)])
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)
    Sm = r.standard_normal((3, 300)).astype(np.float32)
    div = (r.rand(300) + 0.5).astype(np.float32)
    ref = Sm * np.float32(0.25) / div
    ok(pack, pack.alad_scale_scores(p(Sm), 300, 3, 300, p(div), 0.25, None))
    np.testing.assert_allclose(Sm, ref, rtol=1e-6)


# ------------------------------------------------------------------------------------------------- mrsw_bwd.cu
@pytest.fixture(scope="module")
def mrsw_bwd():
    return load_emu("mrsw_bwd")


def run_mrsw_bwd(lib, im, s, nr, nw, G0=None, g0_scale=None, G1=None, sbd_layout=False, region_extent=0):
    """alad_mrsw_scores_bwd on numpy buffers; with sbd_layout the inputs and gradients use the [S,B,d] memory
    layout of alad_model.py:377-378 (viewed as [B,S,d])."""
    from aladin_b200._cabi import MrswBwdArgs
    Bi, S_im, d = im.shape
    Bc, S_s, _ = s.shape

    def lay(x):
        if not sbd_layout:
            x = np.ascontiguousarray(x)
            return x, x.shape[1] * d, d
        x = np.ascontiguousarray(x.transpose(1, 0, 2))            # memory [S,B,d]
        return x, d, x.shape[1] * d

    im_m, im_sb, im_ss = lay(im)
    s_m, s_sb, s_ss = lay(s)
    d_im = np.full(im_m.shape, np.nan, np.float32)
    d_s = np.full(s_m.shape, np.nan, np.float32)
    max_pairs = max(Bi * Bc, 1)
    nbytes = lib.alad_mrsw_bwd_workspace_bytes(Bi, S_im, Bc, S_s, max_pairs)
    ws = workspace(nbytes)
    nr32, nw32 = np.asarray(nr, np.int32), np.asarray(nw, np.int32)
    scale = np.array([g0_scale], np.float32) if g0_scale is not None else None
    a = MrswBwdArgs(im=p(im_m), im_stride_b=im_sb, im_stride_s=im_ss, s=p(s_m), s_stride_b=s_sb, s_stride_s=s_ss,
                    Bi=Bi, S_im=S_im, Bc=Bc, S_s=S_s, d=d, nr=p(nr32), nw=p(nw32),
                    G0=p(G0), ldG0=Bc if G0 is not None else 0, g0_scale=p(scale), G1=p(G1), ldG1=Bc if G1 is not None else 0,
                    d_im=p(d_im), d_s=p(d_s), eps=1e-12, region_extent=region_extent, max_pairs=max_pairs,
                    workspace=p(ws), workspace_bytes=nbytes,
                    d_im_stride_b=im_sb if sbd_layout else 0, d_im_stride_s=im_ss if sbd_layout else 0,
                    d_s_stride_b=s_sb if sbd_layout else 0, d_s_stride_s=s_ss if sbd_layout else 0)
    ok(lib, lib.alad_mrsw_scores_bwd(C.byref(a), None))
    if sbd_layout:
        d_im, d_s = d_im.transpose(1, 0, 2), d_s.transpose(1, 0, 2)
    return d_im, d_s


@pytest.mark.parametrize("shape,sbd", [((6, 7, 9, 12, 64), False), ((3, 4, 60, 80, 32), False), ((4, 5, 8, 11, 20), False)]
                         + big(((5, 4, 35, 53, 64), True)))
def test_mrsw_backward_emulated(mrsw_bwd, shape, sbd):
    """Sparse MrSw backward: the register-tiled pair kernels (d % 32 == 0: <5,2>, <9,3>) and the
    generic kernel (d = 20), contiguous and [S,B,d] gradient layouts, hinge gradient scaled on the device plus a
    dense upstream gradient."""
    from aladin_b200 import synth
    Bi, Bc, S_im, S_s, d = shape
    im, s, il, sl = synth.raw_batch(sum(shape), Bi, Bc, S_im, S_s, d, ragged=True, related=0.5)
    R, W, nr, nw = O.scored_extents(im.shape, s.shape, il, sl)
    r = np.random.RandomState(2)
    G0 = np.zeros((Bi, Bc), np.float32)
    G0[r.rand(Bi, Bc) < 0.4] = 1.0
    G0[0, 0] = -2.0
    G1 = (r.standard_normal((Bi, Bc)) * (r.rand(Bi, Bc) < 0.5)).astype(np.float32)
    d_im, d_s = run_mrsw_bwd(mrsw_bwd, im, s, nr, nw, G0=G0, g0_scale=0.75, G1=G1, sbd_layout=sbd)
    ref_im, ref_s = O.mrsw_backward(im, s, il, sl, 0.75 * G0.astype(np.float64) + G1)
    scale = max(np.abs(ref_im).max(), np.abs(ref_s).max())
    assert np.isfinite(d_im).all() and np.isfinite(d_s).all()
    assert np.abs(d_im - ref_im).max() <= 1e-4 * scale and np.abs(d_s - ref_s).max() <= 1e-4 * scale


def test_mrsw_backward_emulated_reference_golden(mrsw_bwd):
    from conftest import load_golden
    g = load_golden("alignment_loss")
    im = np.ascontiguousarray(np.transpose(g["im_sbd"], (1, 0, 2)))
    s = np.ascontiguousarray(np.transpose(g["s_sbd"], (1, 0, 2)))
    il, sl = g["im_len"].tolist(), g["s_len"].tolist()
    R, W, nr, nw = O.scored_extents(im.shape, s.shape, il, sl)
    G = O.triplet_grad(g["S_mv"], 0.2, True)
    d_im, d_s = run_mrsw_bwd(mrsw_bwd, im, s, nr, nw, G0=np.ascontiguousarray(G, np.float32), sbd_layout=True)
    np.testing.assert_allclose(np.transpose(d_im, (1, 0, 2)), g["dim_mv"], rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(np.transpose(d_s, (1, 0, 2)), g["ds_mv"], rtol=1e-3, atol=2e-6)


# ------------------------------------------------------------------------------------------------- sanitizer
def test_cuda_core_kernels_are_race_free_under_thread_sanitizer():
    """Every test above once more in a child python with gcc's ThreadSanitizer preloaded and the emulated
    libraries built with -fsanitize=thread: CUDA threads are host threads and the barriers are pthread
    barriers, so a shared-memory hand-off without __syncthreads / __syncwarp is reported as a data race."""
    import build_emu
    tsan = build_emu.tsan_runtime()
    if tsan is None:
        pytest.skip("gcc's libtsan.so not found")
    if os.environ.get("ALAD_EMU_TSAN") == "1":
        pytest.skip("already inside the sanitizer run")
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0", ALAD_EMU_TSAN="1")
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-k", "not golden",
                          os.path.abspath(__file__)],
                         env=env, capture_output=True, text=True, timeout=2400, cwd=os.path.dirname(HERE))
    if "FATAL: ThreadSanitizer" in res.stderr:           # the sanitizer runtime cannot start here (e.g. address-space layout)
        pytest.skip("ThreadSanitizer runtime unavailable: " + res.stderr.strip().splitlines()[0][:200])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "ThreadSanitizer: data race" not in res.stderr, res.stderr[:6000]
