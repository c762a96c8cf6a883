"""CPU-only: the CUDA-core kernels of csrc/scan_pool.cu ('scan-sentences', alad/loss.py:136-149) executed by
the host-thread emulator of tests/cuda_emu (the very same source, g++-compiled, one std::thread per CUDA thread)
and checked against the oracle; then the host flow of aladin_b200/scan.py on CPU tensors with the emulated
kernels and torch doubles for the GEMM / normalisation entry points.  This is test infrastructure: it validates
index arithmetic and data flow before the code reaches a GPU; the parity tests proper are the `-m gpu` ones
(tests/test_gpu_zz_scan_sentences.py)."""
import ctypes as C
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from oracle import alad_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")

_I32, _I64, _P = C.c_int32, C.c_int64, C.c_void_p


@pytest.fixture(scope="module")
def emu():
    import build_emu
    from aladin_b200 import _cabi
    lib = C.CDLL(build_emu.build_library())
    for name in ("alad_scan_gram", "alad_scan_gram_bwd", "alad_scan_pool_fwd", "alad_scan_pool_bwd", "alad_scan_apply_pairs"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _cabi.PROTOTYPES[name]
    lib.alad_last_error.restype = C.c_char_p
    return lib


def ptr(a):
    return a.ctypes.data


def unit(x):
    return O.l2_normalize(x).astype(np.float32)


def problem(seed, Bi, Bc, S_im, S_s, d, full_images=False):
    r = np.random.RandomState(seed)
    base = r.standard_normal((max(Bi, Bc), d)).astype(np.float32)
    im = (r.standard_normal((Bi, S_im, d)) + 0.5 * base[:Bi, None]).astype(np.float32)
    s = (r.standard_normal((Bc, S_s, d)) + 0.5 * base[:Bc, None]).astype(np.float32)
    im_len = [S_im] * Bi if full_images else [int(x) for x in r.randint(1, S_im + 1, size=Bi)]
    s_len = [int(x) for x in r.randint(3, S_s + 1, size=Bc)]
    im_len[0], s_len[0] = S_im, S_s
    return im, s, im_len, s_len


def flat_unit_rows(x, extent):
    return np.ascontiguousarray(unit(x)[:, 1:1 + extent].reshape(x.shape[0] * extent, x.shape[2]))


def run_kernels(emu, im, s, im_len, s_len, G=None):
    """S (and, with G, dC / dK) from the emulated kernels; C = unit regions x unit words in numpy."""
    R, W, nr, nw = O.scored_extents(im.shape, s.shape, im_len, s_len)
    Bi, Bc, d = im.shape[0], s.shape[0], im.shape[2]
    xh, yh = flat_unit_rows(im, R), flat_unit_rows(s, W)
    nr32, nw32 = nr.astype(np.int32), nw.astype(np.int32)
    K = np.full((Bc, W, W), np.nan, np.float32)
    assert emu.alad_scan_gram(ptr(yh), Bc, W, d, ptr(nw32), ptr(K), None) == 0, emu.alad_last_error()
    Cm = np.ascontiguousarray((xh.astype(np.float64) @ yh.astype(np.float64).T).astype(np.float32))
    S = np.full((Bi, Bc), 123.0, np.float32)
    rc = emu.alad_scan_pool_fwd(ptr(Cm), Cm.shape[1], Bi, R, Bc, W, ptr(nr32), ptr(nw32), int(nr.max()), int(nw.max()),
                                ptr(K), ptr(S), Bc, None)
    assert rc == 0, emu.alad_last_error()
    out = dict(S=S, K=K, xh=xh, yh=yh, C=Cm, nr=nr, nw=nw, R=R, W=W)
    if G is not None:
        G = np.ascontiguousarray(G, np.float32)
        dC = np.full_like(Cm, np.nan)
        dK = np.zeros_like(K)
        rc = emu.alad_scan_pool_bwd(ptr(Cm), Cm.shape[1], Bi, R, Bc, W, ptr(nr32), ptr(nw32), int(nr.max()), int(nw.max()),
                                    ptr(K), ptr(G), Bc, ptr(dC), dC.shape[1], ptr(dK), None)
        assert rc == 0, emu.alad_last_error()
        d_yh = np.zeros_like(yh)
        assert emu.alad_scan_gram_bwd(ptr(yh), Bc, W, d, ptr(nw32), ptr(dK), ptr(d_yh), None) == 0
        out.update(dC=dC, dK=dK, d_yh_gram=d_yh)
    return out


@pytest.mark.parametrize("shape", [(5, 4, 7, 10, 40), (3, 3, 35, 53, 64), (18, 2, 5, 40, 16), (4, 3, 6, 80, 16),
                                   (5, 4, 7, 10, 40, "smem")])
def test_emulated_kernels_match_oracle(emu, shape, monkeypatch):
    """Forward: register-resident kernel (<= 64 words), shared-memory kernel (77 words, or forced); backward."""
    if shape[-1] == "smem":
        monkeypatch.setenv("ALAD_SCAN_SMEM_FWD", "1")
        shape = shape[:-1]
    Bi, Bc, S_im, S_s, d = shape
    im, s, il, sl = problem(sum(shape), Bi, Bc, S_im, S_s, d)
    r = np.random.RandomState(1)
    G = r.standard_normal((Bi, Bc)).astype(np.float32)
    G[r.rand(Bi, Bc) < 0.3] = 0.0
    k = run_kernels(emu, im, s, il, sl, G)
    nr, nw, R, W = k["nr"], k["nw"], k["R"], k["W"]
    ref = O.scan_scores(im, s, il, sl)
    assert np.array_equal(np.isnan(k["S"]), np.isnan(ref))
    np.testing.assert_allclose(k["S"], ref, rtol=2e-5, atol=2e-6, equal_nan=True)
    # Gram matrices: valid block = Y Y^T, zero outside
    yh3 = k["yh"].reshape(Bc, W, d).astype(np.float64)
    for j in range(Bc):
        Kref = np.zeros((W, W))
        Kref[:nw[j], :nw[j]] = yh3[j, :nw[j]] @ yh3[j, :nw[j]].T
        np.testing.assert_allclose(k["K"][j], Kref, atol=2e-6)
    # per-pair dL/dC and per-caption dL/dK
    C4 = k["C"].reshape(Bi, R, Bc, W).astype(np.float64)
    dC4 = k["dC"].reshape(Bi, R, Bc, W)
    assert np.isfinite(dC4).all()
    for j in range(Bc):
        dK = np.zeros((W, W))
        for i in range(Bi):
            blk = np.zeros((R, W))
            if G[i, j] != 0 and nr[i] and nw[j]:
                dCp, dKp = O._scan_pair_backward(C4[i, :nr[i], j, :nw[j]], k["K"][j, :nw[j], :nw[j]].astype(np.float64),
                                                 float(G[i, j]))
                blk[:nr[i], :nw[j]] = dCp
                dK[:nw[j], :nw[j]] += dKp
            np.testing.assert_allclose(dC4[i, :, j, :], blk, rtol=1e-4, atol=1e-6, err_msg=f"pair {i},{j}")
        np.testing.assert_allclose(k["dK"][j], dK, rtol=1e-4, atol=1e-6)
        ref_g = np.zeros((W, d))
        ref_g[:nw[j]] = 2.0 * dK[:nw[j], :nw[j]] @ yh3[j, :nw[j]]
        np.testing.assert_allclose(k["d_yh_gram"].reshape(Bc, W, d)[j], ref_g, rtol=1e-4, atol=1e-6)


def test_emulated_apply_pairs_equals_dense_products(emu):
    """Pair-list form of d xhat = dC yhat, d yhat = dC' xhat (blocks outside the list are not touched)."""
    r = np.random.RandomState(12)
    Bi, Bc, R, W, d = 6, 5, 7, 9, 300                        # d not a multiple of the 256-column CTA slice
    nr = np.array([7, 0, 3, 7, 5, 1], np.int32)
    nw = np.array([9, 4, 0, 9, 6], np.int32)
    xh = r.standard_normal((Bi * R, d)).astype(np.float32)
    yh = r.standard_normal((Bc * W, d)).astype(np.float32)
    dC = r.standard_normal((Bi * R, Bc * W)).astype(np.float32)
    pairs = np.array([[0, 0], [0, 3], [2, 1], [3, 3], [4, 4], [5, 0], [1, 1], [2, 2]], np.int32)
    d_xh, d_yh = np.zeros_like(xh), np.zeros_like(yh)
    rc = emu.alad_scan_apply_pairs(ptr(dC), Bc * W, ptr(xh), ptr(yh), ptr(pairs), len(pairs), Bi, R, Bc, W, d, ptr(nr), ptr(nw),
                                   int(nr.max()), int(nw.max()), ptr(d_xh), ptr(d_yh), None)
    assert rc == 0, emu.alad_last_error()
    masked = np.zeros_like(dC, dtype=np.float64)
    for i, j in pairs:
        masked[i * R:i * R + nr[i], j * W:j * W + nw[j]] = dC[i * R:i * R + nr[i], j * W:j * W + nw[j]]
    np.testing.assert_allclose(d_xh, masked @ yh.astype(np.float64), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(d_yh, masked.T @ xh.astype(np.float64), rtol=1e-4, atol=1e-5)


def test_emulated_kernels_golden_degenerate_lengths(emu):
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_sentences.npz")))
    k = run_kernels(emu, g["c_im"], g["c_s"], g["c_im_len"].tolist(), g["c_s_len"].tolist())
    assert np.array_equal(np.isnan(k["S"]), np.isnan(g["c_S"]))
    np.testing.assert_allclose(k["S"], g["c_S"], rtol=2e-5, atol=2e-6, equal_nan=True)


def test_emulated_entry_points_reject_bad_arguments(emu):
    z = np.zeros(4, np.float32)
    n = np.zeros(1, np.int32)
    assert emu.alad_scan_pool_fwd(ptr(z), 1, 1, 2, 1, 2, ptr(n), ptr(n), 0, 0, ptr(z), ptr(z), 1, None) == -1   # ldC < Bc*W
    assert b"ldC" in emu.alad_last_error()
    assert emu.alad_scan_pool_fwd(ptr(z), 400, 1, 200, 1, 2, ptr(n), ptr(n), 200, 1, ptr(z), ptr(z), 1, None) == -3
    assert emu.alad_scan_gram(ptr(z), 1, 129, 4, ptr(n), ptr(z), None) == -3


# ----------------------------------------------------------------------------------------------------------
# host flow of aladin_b200/scan.py on CPU tensors: emulated scan kernels + torch doubles for the other entry points
# ----------------------------------------------------------------------------------------------------------
class _Lib:
    def __init__(self, emu):
        self._emu = emu

    def __getattr__(self, name):
        if name.startswith("alad_scan_"):
            return getattr(self._emu, name)
        raise AttributeError(f"{name} has no CPU double")


@pytest.fixture
def scan_on_cpu(emu, monkeypatch):
    from aladin_b200 import _cabi, scan, scoring
    monkeypatch.setattr(_cabi, "lib", lambda: _Lib(emu))
    monkeypatch.setattr(_cabi, "stream_ptr", lambda: None)
    monkeypatch.setattr(_cabi, "check", lambda rc, what: (_ for _ in ()).throw(AssertionError(what)) if rc else None)
    monkeypatch.setattr(scoring, "_to_dev_group", lambda arrs, dev: [torch.from_numpy(np.ascontiguousarray(a)) for a in arrs])
    monkeypatch.setattr(scoring, "unit_rows", lambda x, eps=0.0: torch.nn.functional.normalize(x, dim=1, eps=eps))

    def dot_scores(im, s, precision=None, normalize=False, eps=0.0, out=None):
        assert im.is_contiguous() and s.is_contiguous() and not normalize
        res = (im.double() @ s.double().t()).float()
        if out is not None:
            assert out.shape == res.shape and out.stride(1) == 1
            out.copy_(res)
            return out
        return res

    def normalize_bwd_(x, dx):
        xd, g = x.double(), dx.double()
        n = xd.norm(dim=1, keepdim=True).clamp_min(1e-12)
        xh = xd / n
        dx.copy_(((g - xh * (xh * g).sum(1, keepdim=True)) / n).float())

    monkeypatch.setattr(scoring, "dot_scores", dot_scores)
    monkeypatch.setattr(scan, "_normalize_bwd_", normalize_bwd_)
    return scan


@pytest.mark.parametrize("chunk_images,sparse", [(None, False), (2, False), (None, True), (2, True)])
def test_scan_host_flow_on_cpu(scan_on_cpu, monkeypatch, chunk_images, sparse):
    from aladin_b200 import scoring
    scan = scan_on_cpu
    if sparse:
        monkeypatch.setattr(scan, "SPARSE_FRACTION", 1)       # pair-list backward whatever the density of dL/dS
    Bi, Bc, S_im, S_s, d = 5, 4, 7, 10, 24
    im, s, il, sl = problem(11, Bi, Bc, S_im, S_s, d)
    if chunk_images:
        monkeypatch.setattr(scan, "_CHUNK_BYTES", (S_im - 1) * Bc * (S_s - 3) * 4 * chunk_images)
    # the training layout: [S,B,d] tensors viewed as [B,S,d] (alad_model.py:377-378)
    im_t = torch.from_numpy(np.ascontiguousarray(im.transpose(1, 0, 2))).permute(1, 0, 2)
    s_t = torch.from_numpy(s)
    counts = scoring.scored_counts(im_t.shape, s_t.shape, il, sl)
    S = scan.scan_scores(im_t, s_t, counts)
    np.testing.assert_allclose(S.numpy(), O.scan_scores(im, s, il, sl), rtol=2e-5, atol=2e-6, equal_nan=True)
    G = np.random.RandomState(2).standard_normal((Bi, Bc)).astype(np.float32)
    G[1, 2] = 0.0
    if sparse:
        G[np.random.RandomState(3).rand(Bi, Bc) < 0.6] = 0.0
    d_im, d_s = scan.scan_backward(im_t, s_t, counts, torch.from_numpy(G))
    ref_im, ref_s = O.scan_backward(im, s, il, sl, G)
    np.testing.assert_allclose(d_im.numpy(), ref_im, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(d_s.numpy(), ref_s, rtol=1e-4, atol=2e-6)


def test_scan_host_flow_reference_gradient_full_length(scan_on_cpu):
    """End to end against the unmodified reference's autograd (finite: every image full length)."""
    from aladin_b200 import scoring
    scan = scan_on_cpu
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_sentences.npz")))
    il, sl = g["d_im_len"].tolist(), g["d_s_len"].tolist()
    im_t, s_t = torch.from_numpy(g["d_im"]), torch.from_numpy(g["d_s"])
    counts = scoring.scored_counts(im_t.shape, s_t.shape, il, sl)
    np.testing.assert_allclose(scan.scan_scores(im_t, s_t, counts).numpy(), g["d_S"], rtol=2e-5, atol=2e-6)
    d_im, d_s = scan.scan_backward(im_t, s_t, counts, torch.from_numpy(g["d_Gup"]))
    np.testing.assert_allclose(d_im.numpy(), g["d_dim"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(d_s.numpy(), g["d_ds"], rtol=1e-4, atol=2e-6)


def test_scan_empty_word_extent_raises_like_reference(scan_on_cpu):
    from aladin_b200 import scoring
    im_t, s_t = torch.zeros(2, 4, 8), torch.zeros(2, 3, 8)
    counts = scoring.scored_counts(im_t.shape, s_t.shape, [4, 4], [3, 3])
    with pytest.raises(IndexError):
        scan_on_cpu.scan_scores(im_t, s_t, counts)


def test_scan_kernels_are_race_free_under_thread_sanitizer():
    """Every shared-memory hand-off between CUDA threads needs a barrier: the emulated kernels run under
    ThreadSanitizer (host threads + pthread barriers) and must not report a data race."""
    import subprocess
    import build_emu
    tsan = build_emu.tsan_runtime()
    if tsan is None:
        pytest.skip("gcc's libtsan.so not found")
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu", "tsan_scan.py")
    res = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
    if "FATAL: ThreadSanitizer" in res.stderr:           # the sanitizer runtime cannot start here (e.g. address-space layout)
        pytest.skip("ThreadSanitizer runtime unavailable: " + res.stderr.strip().splitlines()[0][:200])
    assert "scan tsan ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ThreadSanitizer: data race" not in res.stderr, res.stderr[:4000]
