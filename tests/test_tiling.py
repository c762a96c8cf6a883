"""Host bookkeeping: the native greedy tiling helper (alad_region_tiles, host-only C) against its numpy
restatement, the metadata upload grouping, and the reference's length arithmetic (alad/loss.py:87-116)."""
import numpy as np
import pytest

from aladin_b200 import tiling


@pytest.mark.parametrize("mode", ["ragged", "uniform", "wide", "narrow"])
def test_region_tiles_native_matches_numpy(mode):
    for trial in range(40):
        r = np.random.RandomState(1000 * trial + len(mode))
        Ni = int(r.randint(0, 700))
        if mode == "ragged":
            nr = r.randint(0, 35, Ni)
        elif mode == "uniform":
            nr = np.full(Ni, 34)
        elif mode == "wide":
            nr = (r.rand(Ni) < 0.7) * r.randint(1, tiling.TILE_N + 1, Ni)
        else:
            nr = r.randint(1, 8, Ni)           # > MAX_SEG images would fit the columns: the image cap closes tiles
        clamp = r.rand(Ni) < 0.5
        off_a, tab_a, rows_a = tiling.build_region_tiles(nr, clamp)
        off_b, tab_b, rows_b = tiling.build_region_tiles_numpy(nr, clamp)
        assert rows_a == rows_b == int(np.sum(nr))
        assert np.array_equal(off_a, off_b)
        assert tab_a.shape == tab_b.shape and np.array_equal(tab_a, tab_b)
        if len(tab_a):
            assert int(tab_a[:, 2].max()) <= tiling.MAX_SEG


def test_region_tiles_rejects_oversized_image():
    with pytest.raises(ValueError):
        tiling.build_region_tiles(np.array([3, tiling.TILE_N + 1]), np.array([False, False]))


def test_valid_counts_python_slice_semantics():
    # l = len - drop; mask[l:] = True on `extent` slots: negative l counts from the end (alad/loss.py:103-112)
    got = tiling.valid_counts([0, 1, 2, 3, 10, 60], 3, 50)
    assert got.tolist() == [47, 48, 49, 0, 7, 50]
