"""Host-side bookkeeping for the packed layout: valid token counts (the reference's
length arithmetic, alad/loss.py:87-116), packed row offsets and the 240-column region
tile table consumed by ``alad_mrsw_scores_fwd``.  Pure numpy; no device work."""
import numpy as np

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
    """Greedy tiling of consecutive images into tiles of <= TILE_N packed region rows and
    <= MAX_SEG images.  Images with no valid region own no column (their score row stays 0).

    Returns (row_off[int64 Ni], table[uint32 T, 12], n_rows).  Table row layout = struct
    alad_ntile: row_start, img0, nseg, clamp_bits, start_mask[8] (bit ncols = sentinel)."""
    nr = np.asarray(nr, dtype=np.int64)
    clamp = np.asarray(clamp, dtype=bool)
    if nr.size and int(nr.max()) > TILE_N:
        raise ValueError(f"an image has {int(nr.max())} scored regions; the kernel supports at most {TILE_N}")
    row_off, n_rows = exclusive_cumsum(nr)
    tiles = []
    Ni = len(nr)
    uniform = Ni > 0 and int(nr.min()) == int(nr.max()) and nr[0] > 0
    if uniform:
        # closed form: every tile holds `per` images of `w` columns
        w = int(nr[0])
        per = max(1, min(MAX_SEG, TILE_N // w))
        n_t = (Ni + per - 1) // per
        table = np.zeros((n_t, NTILE_WORDS), dtype=np.uint32)
        img0 = np.arange(n_t, dtype=np.int64) * per
        nseg = np.minimum(per, Ni - img0)
        table[:, 0] = (img0 * w).astype(np.uint32)
        table[:, 1] = img0.astype(np.uint32)
        table[:, 2] = nseg.astype(np.uint32)
        cb = np.zeros(n_t, dtype=np.uint64)
        cl = clamp.astype(np.uint64)
        for s in range(per):
            idx = img0 + s
            ok = idx < Ni
            cb[ok] |= cl[idx[ok]] << np.uint64(s)
        table[:, 3] = cb.astype(np.uint32)
        for t_nseg in np.unique(nseg):
            mask = np.zeros(8, dtype=np.uint32)
            for s in range(int(t_nseg) + 1):           # segment starts + sentinel
                c = s * w
                mask[c >> 5] |= np.uint32(1 << (c & 31))
            table[nseg == t_nseg, 4:12] = mask
        return row_off, table, n_rows
    i = 0
    while i < Ni:
        if nr[i] == 0:
            i += 1
            continue
        rec = np.zeros(NTILE_WORDS, dtype=np.uint32)
        rec[0] = row_off[i]
        rec[1] = i
        cols = 0
        seg = 0
        while i < Ni and nr[i] > 0 and seg < MAX_SEG and cols + nr[i] <= TILE_N:
            rec[4 + (cols >> 5)] |= np.uint32(1 << (cols & 31))
            if clamp[i]:
                rec[3] |= np.uint32(1 << seg)
            cols += int(nr[i])
            seg += 1
            i += 1
        rec[4 + (cols >> 5)] |= np.uint32(1 << (cols & 31))      # sentinel at column `cols`
        rec[2] = seg
        tiles.append(rec)
    table = np.stack(tiles) if tiles else np.zeros((0, NTILE_WORDS), dtype=np.uint32)
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
    return ((n_rows + TILE_M - 1) // TILE_M) * TILE_M


def round_up(x, m):
    return ((x + m - 1) // m) * m
