"""TEST INFRASTRUCTURE ONLY: run the emulated kernels of csrc/scan_pool.cu under ThreadSanitizer.  Must be started
with gcc's libtsan.so in LD_PRELOAD (tests/test_scan_emu.py does that); prints "scan tsan ok" when the results
match the oracle -- data races are reported by the sanitizer on stderr."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(os.path.dirname(HERE))]

import build_emu  # noqa: E402
from aladin_b200 import _cabi  # noqa: E402
from oracle import alad_oracle as O  # noqa: E402

lib = C.CDLL(build_emu.build_library(tsan=True))
for name in ("alad_scan_gram", "alad_scan_gram_bwd", "alad_scan_pool_fwd", "alad_scan_pool_bwd", "alad_scan_apply_pairs"):
    getattr(lib, name).restype, getattr(lib, name).argtypes = _cabi.PROTOTYPES[name]

r = np.random.RandomState(0)
Bi, Bc, S_im, S_s, d = 9, 2, 6, 40, 16                     # > 32 words: lanes own two columns; 9 images: 3 rounds of 4 warps
im = r.standard_normal((Bi, S_im, d)).astype(np.float32)
s = r.standard_normal((Bc, S_s, d)).astype(np.float32)
il, sl = [6, 3, 6, 2, 5, 6, 4, 6, 6], [40, 20]
R, W, nr, nw = O.scored_extents(im.shape, s.shape, il, sl)
xh = np.ascontiguousarray(O.l2_normalize(im)[:, 1:1 + R].reshape(Bi * R, d))
yh = np.ascontiguousarray(O.l2_normalize(s)[:, 1:1 + W].reshape(Bc * W, d))
nr32, nw32 = nr.astype(np.int32), nw.astype(np.int32)


def p(a):
    return a.ctypes.data


K = np.zeros((Bc, W, W), np.float32)
assert lib.alad_scan_gram(p(yh), Bc, W, d, p(nw32), p(K), None) == 0
Cm = np.ascontiguousarray(xh @ yh.T)
S = np.zeros((Bi, Bc), np.float32)
for smem_kernel in (False, True):                          # register-resident forward kernel, then the shared-memory one
    if smem_kernel:
        os.environ["ALAD_SCAN_SMEM_FWD"] = "1"
    S[:] = 0
    assert lib.alad_scan_pool_fwd(p(Cm), Cm.shape[1], Bi, R, Bc, W, p(nr32), p(nw32), int(nr.max()), int(nw.max()), p(K),
                                  p(S), Bc, None) == 0
    assert np.abs(S - O.scan_scores(im, s, il, sl)).max() < 1e-5
G = r.standard_normal((Bi, Bc)).astype(np.float32)
dC, dK, dy = np.zeros_like(Cm), np.zeros_like(K), np.zeros_like(yh)
assert lib.alad_scan_pool_bwd(p(Cm), Cm.shape[1], Bi, R, Bc, W, p(nr32), p(nw32), int(nr.max()), int(nw.max()), p(K), p(G),
                              Bc, p(dC), dC.shape[1], p(dK), None) == 0
assert lib.alad_scan_gram_bwd(p(yh), Bc, W, d, p(nw32), p(dK), p(dy), None) == 0
pairs = np.array([[0, 0], [2, 1], [8, 0], [8, 1]], np.int32)
d_xh = np.zeros_like(xh)
assert lib.alad_scan_apply_pairs(p(dC), dC.shape[1], p(xh), p(yh), p(pairs), len(pairs), Bi, R, Bc, W, d, p(nr32), p(nw32),
                                 int(nr.max()), int(nw.max()), p(d_xh), p(dy), None) == 0
assert np.abs(S - O.scan_scores(im, s, il, sl)).max() < 1e-5
print("scan tsan ok")
