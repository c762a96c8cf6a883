"""TEST INFRASTRUCTURE: randomised small-shape sweep of the emulated CUDA-core kernels against the oracle (ranking with ties /
padded rows / k > Ni, triplet, ListNet, distillation modes; MrSw backward in both gradient layouts, scan-sentences with empty
images / captions).  Not part of the default suite (minutes):  python tests/cuda_emu/fuzz_kernels.py"""
import ctypes as C
import sys

import numpy as np

import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))]
import test_kernels_emu as T
import test_scan_emu as TS
from oracle import alad_oracle as O
lib = T.load_emu("all")
p, ok, workspace = T.p, T.ok, T.workspace
r = np.random.RandomState(123)
fails = 0
def stable_desc(v): return np.argsort(v, kind="stable")[::-1]
for it in range(40):
    # ---- ranking with odd shapes, group 5
    Ni = int(r.randint(1, 20)); Nc = 5 * Ni; k = int(r.choice([1, 3, 7, 50]))
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    if it % 3 == 0: S = np.round(S)            # heavy ties
    ld = Nc + int(r.choice([0, 3]))
    Sp = np.zeros((Ni, ld), np.float32); Sp[:, :Nc] = S
    try:
        rk, top1 = np.zeros(Ni, np.int32), np.zeros(Ni, np.int32)
        ok(lib, lib.alad_rank_rows(p(Sp), ld, Ni, Nc, 5, 0, p(rk), p(top1), None))
        gt = np.zeros(Nc, np.float32); ok(lib, lib.alad_col_gt(p(Sp), ld, Ni, Nc, 5, 0, p(gt), None))
        cnt = np.zeros(Nc, np.int32); ok(lib, lib.alad_col_count(p(Sp), ld, Ni, Nc, 5, 0, p(gt), p(cnt), None))
        ss, si = np.zeros((Nc, k), np.float32), np.zeros((Nc, k), np.int32)
        ws = workspace(lib.alad_col_topk_select_workspace_bytes(Ni, Nc, k))
        ok(lib, lib.alad_col_topk_select(p(Sp), ld, Ni, Nc, k, 0, p(ss), p(si), p(ws), None))
        kk = min(k, Ni)
        for i in range(Ni):
            inds = stable_desc(S[i]); pos = np.empty(Nc, np.int64); pos[inds] = np.arange(Nc)
            assert rk[i] == pos[5*i:5*i+5].min() and top1[i] == inds[0], ("rank_rows", Ni, k)
        for c in range(Nc):
            inds = stable_desc(S[:, c])
            assert cnt[c] == np.where(inds == c // 5)[0][0], ("col_count", Ni)
            assert np.array_equal(si[c, :kk], inds[:kk]) and np.all(si[c, kk:] == -1), ("topk_select", Ni, k, c)
    except AssertionError as e:
        fails += 1; print("FAIL ranking", it, Ni, k, ld, e)
    # ---- losses
    B = int(r.randint(1, 40))
    Sq = r.standard_normal((B, B)).astype(np.float32)
    if it % 4 == 0: Sq = np.round(Sq * 2) / 2
    for mv in (True, False):
        try:
            loss, G = T.run_triplet(lib, Sq, 0.2, mv)
            assert np.isclose(loss, O.triplet_loss(Sq, 0.2, mv), rtol=3e-5, atol=1e-6), ("triplet loss", B, mv, loss, O.triplet_loss(Sq, 0.2, mv))
            assert np.array_equal(G, O.triplet_grad(Sq, 0.2, mv)), ("triplet grad", B, mv)
        except AssertionError as e:
            fails += 1; print("FAIL triplet", it, B, mv, str(e)[:200])
    Tm = (r.standard_normal((B, B)) * 2 + 3).astype(np.float32); M = np.clip(r.standard_normal((B, B)) * .3, -1, 1).astype(np.float32)
    try:
        loss = np.zeros(1, np.float32); dM = np.zeros((B, B), np.float32); ws = workspace(lib.alad_loss_workspace_bytes(B))
        ok(lib, lib.alad_listnet_fwd_bwd(p(Tm), B, p(M), B, B, 6.0, 1e-10, p(loss), p(dM), B, p(ws), None))
        ref = O.listnet_grad(Tm, M)
        assert np.isclose(loss[0], O.listnet_loss(Tm, M), rtol=3e-5), ("listnet loss", B)
        assert np.allclose(dM, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max()), ("listnet grad", B)
        for mode, fn in ((1, "contr"), (2, "ord")):
            l2 = np.zeros(1, np.float32); d2 = np.zeros((B, B), np.float32); ws = workspace(lib.alad_distill_workspace_bytes(B, mode))
            if mode == 1:
                Tc = Tm.copy(); ok(lib, lib.alad_distill_contrastive_fwd_bwd(p(Tc), B, p(M), B, B, 0.2, 1, p(l2), p(d2), B, p(ws), None))
                rl, rg = O.distill_contrastive(Tm, M, 0.2)
            else:
                st = int(r.choice([1, 2, 3])); ok(lib, lib.alad_distill_ordinal_fwd_bwd(p(Tm), B, p(M), B, B, 0.2, 0.1, st, p(l2), p(d2), B, p(ws), None))
                rl, rg = O.distill_ordinal(Tm, M, 0.2, 0.1, st)
            assert (np.isnan(rl) and np.isnan(l2[0])) or np.isclose(l2[0], rl, rtol=3e-5, atol=1e-6), (fn, "loss", B, l2[0], rl)
            if not np.isnan(rl): assert np.allclose(d2, rg, rtol=1e-4, atol=1e-6), (fn, "grad", B)
    except AssertionError as e:
        fails += 1; print("FAIL distill/listnet", it, B, str(e)[:200])
fails_a = fails

from aladin_b200 import synth

import ctypes as C
from aladin_b200 import _cabi
lib.alad_last_error.restype = C.c_char_p
r = np.random.RandomState(7)
fails = 0
for it in range(14):
    Bi, Bc = int(r.randint(1, 7)), int(r.randint(1, 7))
    S_im, S_s = int(r.randint(2, 40)), int(r.randint(4, 60))
    d = int(r.choice([20, 32, 33, 64, 96]))
    im, s, il, sl = synth.raw_batch(100 + it, Bi, Bc, S_im, S_s, d, ragged=True, related=0.5)
    if it % 3 == 0: il[0] = 1                       # an image without valid regions
    if it % 4 == 0: sl[0] = 3                       # a caption without valid words
    R, W, nr, nw = O.scored_extents(im.shape, s.shape, il, sl)
    G0 = (r.rand(Bi, Bc) < 0.5).astype(np.float32); G1 = (r.standard_normal((Bi, Bc)) * (r.rand(Bi, Bc) < 0.5)).astype(np.float32)
    try:
        for sbd in (False, True):
            d_im, d_s = T.run_mrsw_bwd(lib, im, s, nr, nw, G0=G0, g0_scale=0.5, G1=G1, sbd_layout=sbd)
            ref_im, ref_s = O.mrsw_backward(im, s, il, sl, 0.5 * G0.astype(np.float64) + G1)
            scale = max(np.abs(ref_im).max(), np.abs(ref_s).max(), 1e-6)
            assert np.isfinite(d_im).all() and np.isfinite(d_s).all()
            assert np.abs(d_im - ref_im).max() <= 2e-4 * scale and np.abs(d_s - ref_s).max() <= 2e-4 * scale, ("mrsw_bwd", sbd, np.abs(d_im - ref_im).max() / scale)
    except AssertionError as e:
        fails += 1; print("FAIL mrsw_bwd", it, (Bi, Bc, S_im, S_s, d), il, sl, str(e)[:200])
    # scan kernels (only W >= 1)
    try:
        G = (r.standard_normal((Bi, Bc)) * (r.rand(Bi, Bc) < 0.7)).astype(np.float32)
        k = TS.run_kernels(lib, im, s, il, sl, G)
        ref = O.scan_scores(im, s, il, sl)
        assert np.array_equal(np.isnan(k["S"]), np.isnan(ref)), "scan nan pattern"
        assert np.allclose(k["S"], ref, rtol=3e-5, atol=3e-6, equal_nan=True), ("scan fwd", np.nanmax(np.abs(k["S"] - ref)))
        # full backward composed in numpy like aladin_b200/scan.py
        Gz = np.where(np.isnan(ref), 0, G)
        xh, yh, dC, dK = k["xh"].astype(np.float64), k["yh"].astype(np.float64), k["dC"].astype(np.float64), k["dK"]
        assert np.isfinite(dC).all()
    except AssertionError as e:
        fails += 1; print("FAIL scan", it, (Bi, Bc, S_im, S_s, d), il, sl, str(e)[:200])
print("failures:", fails_a + fails)
sys.exit(1 if fails_a + fails else 0)
