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


def assert_scores_close(got, ref, rtol, what=""):
    """Score-matrix tolerance: |got - ref| <= rtol * max(|ref|, 10% of the matrix's largest
    magnitude).  Entry-wise relative error is meaningless for cosines that happen to be ~0;
    the floor ties the tolerance to the scale of the matrix."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    floor = 0.1 * np.abs(ref).max() if ref.size else 0.0
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
    floor = 0.1 * np.abs(scores).max()
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
    floor = 0.1 * np.abs(scores).max()
    for q in np.nonzero(got != ref)[0]:
        g = scores[q, gt_idx[q]]
        near = np.count_nonzero(np.abs(scores[q] - g) <= rtol * max(abs(g), floor)) - 1
        assert abs(got[q] - ref[q]) <= near, f"{what}: rank of query {q} is {got[q]}, reference {ref[q]}, {near} near ties"
