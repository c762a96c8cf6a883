"""Host-side bookkeeping for the packed layout: valid token counts (the reference's
length arithmetic, alad/loss.py:87-116), packed row offsets and the 240-column region
tile table consumed by ``alad_mrsw_scores_fwd``.  No device work: numpy plus the host-side
``alad_region_tiles`` helper of the C ABI (a sequential greedy loop is 400 us in Python for a
batch of 512 images -- half of a training step's forward -- and ~1 us in C)."""
import numpy as np

from . import _cabi
from ._cabi import MAX_SEG, NTILE_WORDS, TILE_M, TILE_N


def valid_counts(lengths, drop, extent):
    """Unmasked leading slots per item after ``l = len - drop; mask[l:] = True`` on a row of
    ``extent`` slots (alad/loss.py:89-90,103-112), Python slice semantics included."""
    l = np.asarray(lengths, dtype=np.int64).reshape(-1) - int(drop)
    out = np.where(l >= 0, np.minimum(l, extent), np.maximum(extent + l, 0))
    return out.astype(np.int32)


def exclusive_cumsum(counts):
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return off[:-1], int(off[-1])


def build_region_tiles(nr, clamp):
    """Greedy tiling of consecutive images into tiles of <= TILE_N packed region rows and <= MAX_SEG
    images (images with no valid region own no column) through ``alad_region_tiles``.
    Returns (row_off[int64 Ni], table[uint32 T, 20], n_rows); table row layout = struct alad_ntile."""
    nr = np.ascontiguousarray(nr, dtype=np.int32)
    clamp = np.ascontiguousarray(clamp, dtype=np.uint8)
    Ni = len(nr)
    if Ni and int(nr.max()) > TILE_N:
        raise ValueError(f"an image has {int(nr.max())} scored regions; the kernel supports at most {TILE_N}")
    table = np.empty((max(Ni, 1), NTILE_WORDS), dtype=np.uint32)
    row_off = np.empty(Ni, dtype=np.int64)
    n_t = _cabi.lib().alad_region_tiles(nr.ctypes.data, clamp.ctypes.data, Ni, table.ctypes.data, len(table),
                                        row_off.ctypes.data)
    _cabi.check(min(n_t, 0), "alad_region_tiles")
    n_rows = int(row_off[-1]) + int(nr[-1]) if Ni else 0
    return row_off, table[:n_t], n_rows


def build_region_tiles_numpy(nr, clamp):
    """numpy restatement of ``alad_region_tiles`` (cross-check in tests/test_tiling.py).  Greedy tiling of consecutive images into tiles of <= TILE_N packed region rows and
    <= MAX_SEG images.  Images with no valid region own no column (their score row stays 0).

    Returns (row_off[int64 Ni], table[uint32 T, 20], n_rows).  Table row layout = struct
    alad_ntile: row_start, img0, nseg, clamp_bits, seg[32] (uint16: first column | width << 8)."""
    nr = np.asarray(nr, dtype=np.int64)
    clamp = np.asarray(clamp, dtype=bool)
    if nr.size and int(nr.max()) > TILE_N:
        raise ValueError(f"an image has {int(nr.max())} scored regions; the kernel supports at most {TILE_N}")
    row_off, n_rows = exclusive_cumsum(nr)
    Ni = len(nr)
    # tile id of every image (greedy, sequential): images without regions close the current tile
    tile_of = np.full(Ni, -1, dtype=np.int64)
    col_of = np.zeros(Ni, dtype=np.int64)
    seg_of = np.zeros(Ni, dtype=np.int64)
    uniform = Ni > 0 and int(nr.min()) == int(nr.max()) and nr[0] > 0
    if uniform:
        w = int(nr[0])
        per = max(1, min(MAX_SEG, TILE_N // w))
        idx = np.arange(Ni, dtype=np.int64)
        tile_of = idx // per
        seg_of = idx % per
        col_of = seg_of * w
        n_t = int(tile_of[-1]) + 1
    else:
        t = -1
        cols = seg = 0
        open_tile = False
        for i in range(Ni):
            n = int(nr[i])
            if n == 0:
                open_tile = False
                continue
            if not open_tile or seg >= MAX_SEG or cols + n > TILE_N:
                t += 1
                cols = seg = 0
                open_tile = True
            tile_of[i], col_of[i], seg_of[i] = t, cols, seg
            cols += n
            seg += 1
        n_t = t + 1
    table = np.zeros((n_t, NTILE_WORDS), dtype=np.uint32)
    if n_t == 0:
        return row_off, table, n_rows
    has = tile_of >= 0
    ii = np.nonzero(has)[0]
    tt, ss = tile_of[ii], seg_of[ii]
    first = ss == 0
    table[tt[first], 0] = row_off[ii[first]].astype(np.uint32)
    table[tt[first], 1] = ii[first].astype(np.uint32)
    np.add.at(table[:, 2], tt, 1)
    np.bitwise_or.at(table[:, 3], tt, (clamp[ii].astype(np.uint32) << ss.astype(np.uint32)))
    seg16 = (col_of[ii] | (nr[ii] << 8)).astype(np.uint32)
    np.bitwise_or.at(table, (tt, 4 + ss // 2), seg16 << (16 * (ss % 2)).astype(np.uint32))
    return row_off, table, n_rows


def gemm_tiles(n_rows):
    """Tile table for the plain-GEMM epilogue: consecutive blocks of TILE_N rows."""
    n_t = (n_rows + TILE_N - 1) // TILE_N
    table = np.zeros((n_t, NTILE_WORDS), dtype=np.uint32)
    table[:, 0] = np.arange(n_t, dtype=np.uint32) * TILE_N
    table[:, 1] = table[:, 0]
    table[:, 2] = 1
    return table


def padded_rows(n_rows):
    """row_cap length: word rows rounded up to a CTA pair's 2 x TILE_M rows."""
    return ((n_rows + 2 * TILE_M - 1) // (2 * TILE_M)) * (2 * TILE_M)


def round_up(x, m):
    return ((x + m - 1) // m) * m
