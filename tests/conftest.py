import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) on a GPU-less box only when explicitly selected."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SCORE_FLOOR = 0.01
# Matrices whose entries are CANCELLING sums (plain dot products of global vectors, 'sum' / 'mean' pooling of signed
# cosines): an fp32 dot product -- the reference's own torch.mm included -- is accurate relative to sum_k |a_k b_k|, not
# relative to a result that happens to cancel to ~0, and for these matrices that scale is >= 10 % of the largest entry.
# With the 1 % floor the worst entries of exactly these matrices sit at 1.0e-4 .. 2.4e-4 (round 2 measurement); the
# max-pooled scores of the hot path (sums of non-negative maxima, no cancellation) pass with 1 %.
CANCELLING_FLOOR = 0.1


def assert_scores_close(got, ref, rtol, what="", floor=SCORE_FLOOR):
    """Score-matrix tolerance: |got - ref| <= rtol * max(|ref|, floor * the matrix's largest magnitude), floor = 1 %
    (round 1 used 10 % everywhere: VERDICT r1, weak 3).  Entry-wise relative error is meaningless for scores that happen
    to be ~0; the floor ties the tolerance to the scale of the matrix."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    floor = floor * np.abs(ref).max() if ref.size else 0.0
    err = np.abs(got - ref) / np.maximum(np.abs(ref), max(floor, 1e-30))
    assert err.max() <= rtol, f"{what}: max scaled error {err.max():.3e} > {rtol:.1e}"


def assert_order_equal_up_to_ties(got, ref, scores, rtol, what=""):
    """Top-k index lists must be identical except where the two candidates' reference scores
    differ by less than the score tolerance (BASELINE.json: "top-k indices must be bit-exact
    except for ties inside that tolerance").  got/ref: [Q,k] indices, scores: [Q,n] reference."""
    got = np.asarray(got).astype(np.int64)
    ref = np.asarray(ref).astype(np.int64)
    scores = np.asarray(scores, dtype=np.float64)
    assert got.shape == ref.shape, what
    q, p = np.nonzero(got != ref)
    if q.size == 0:
        return
    floor = SCORE_FLOOR * np.abs(scores).max()
    a, b = scores[q, got[q, p]], scores[q, ref[q, p]]
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    assert err.max() <= rtol, f"{what}: {q.size} order differences, worst score gap {err.max():.3e} > {rtol:.1e}"
    for qq in np.unique(q):                      # and the sets may only differ by near-tied members
        extra = np.setxor1d(got[qq], ref[qq])
        if extra.size:
            kth = scores[qq, ref[qq, -1]]
            gap = np.abs(scores[qq, extra] - kth) / max(abs(kth), floor)
            assert gap.max() <= rtol, f"{what}: top-k sets differ beyond ties for query {qq}"


def assert_ranks_equal_up_to_ties(got, ref, scores, gt_idx, rtol, what=""):
    """Ranks must be identical except where competitors sit within the score tolerance of the
    ground truth: |got - ref| may not exceed the number of such near-tied competitors."""
    got, ref = np.asarray(got, dtype=np.int64), np.asarray(ref, dtype=np.int64)
    scores = np.asarray(scores, dtype=np.float64)
    floor = SCORE_FLOOR * np.abs(scores).max()
    for q in np.nonzero(got != ref)[0]:
        g = scores[q, gt_idx[q]]
        near = np.count_nonzero(np.abs(scores[q] - g) <= rtol * max(abs(g), floor)) - 1
        assert abs(got[q] - ref[q]) <= near, f"{what}: rank of query {q} is {got[q]}, reference {ref[q]}, {near} near ties"


# ----------------------------------------------------------------------------------------------------------
# "virtual B200": the unchanged host layer and the unchanged GPU test bodies on CPU tensors
# ----------------------------------------------------------------------------------------------------------
@pytest.fixture
def virtual_b200(monkeypatch):
    """TEST INFRASTRUCTURE.  Routes aladin_b200 to the emulated library of tests/cuda_emu (every CUDA-core kernel
    compiled from its unchanged .cu source and executed by host threads, plus a CPU double for the tcgen05 entry
    point) and makes "cuda" mean "cpu" for the duration of one test, so that selected `-m gpu` test functions can be
    called as they are on the GPU-less box.  Small sizes only: one std::thread per CUDA thread."""
    import ctypes
    import shutil
    import torch
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
    import build_emu
    from aladin_b200 import _cabi, scoring
    lib = ctypes.CDLL(build_emu.build_library())
    for name, (res, args) in _cabi.PROTOTYPES.items():
        fn = getattr(lib, name)                  # the emulated library exports the complete C ABI
        fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_cabi, "_lib", lib)
    monkeypatch.setattr(_cabi, "stream_ptr", lambda: None)
    monkeypatch.setattr(scoring, "_cuda_ok", lambda: True)
    monkeypatch.setattr(scoring, "_META_CACHE", type(scoring._META_CACHE)())

    def to_dev(arr, device):
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        return torch.from_numpy(arr.copy())

    monkeypatch.setattr(scoring, "_to_dev", to_dev)

    def cpu_device(dev):
        return "cpu" if dev is not None and str(getattr(dev, "type", dev)).startswith("cuda") else dev

    orig_empty = torch.empty
    for name in ("tensor", "empty", "zeros", "ones", "full", "arange", "empty_like", "zeros_like", "randn", "rand", "eye"):
        orig = getattr(torch, name)

        def factory(*a, _orig=orig, **k):
            if "device" in k:
                k["device"] = cpu_device(k["device"])
            k.pop("pin_memory", None)
            t = _orig(*a, **k)
            if t.numel() and t.data_ptr() % 256:
                # device allocations are 256-byte aligned (the C ABI checks its workspaces); CPU ones are not
                nbytes = t.numel() * t.element_size()
                raw = orig_empty(nbytes + 256, dtype=torch.uint8)
                off = (-raw.data_ptr()) % 256
                al = raw[off:off + nbytes].view(t.dtype).view(t.shape)
                al.copy_(t.detach())
                t = al.requires_grad_(t.requires_grad)
            return t

        monkeypatch.setattr(torch, name, factory)
    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(cpu_device(x) if isinstance(x, (str, torch.device)) else x for x in a)
        if "device" in k:
            k["device"] = cpu_device(k["device"])
        return orig_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", to)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)

    # streams and events: everything runs synchronously on the calling thread
    class Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_stream(self, s):
            pass

        def wait_event(self, e):
            pass

        def synchronize(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    class Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a, **k):
            pass

        def wait(self, *a, **k):
            pass

        def synchronize(self):
            pass

        def query(self):
            return True

        def elapsed_time(self, other):
            return 0.0

    monkeypatch.setattr(torch.cuda, "Stream", Stream)
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: s if s is not None else Stream())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 1)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return lib
