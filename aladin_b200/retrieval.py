"""All-pairs retrieval engine: upload -> pack -> fused MrSw scores -> exact ranks + top-k.

This is the fused replacement for the per-query Python loops of alad/evaluation.py:175-223
(i2t) and :263-313 (t2i): the gallery goes to the device once, one pass over the
Ni x Nc score block serves both directions, ranks come from "count ahead of the ground
truth" kernels instead of per-query numpy argsorts.

Multi-GPU: gallery images are split into contiguous image blocks, captions are
replicated.  i2t needs no exchange.  t2i exchanges, through torch.distributed / NCCL, ONE all-gather that
carries the ground-truth scores, the per-shard top-k candidates and the i2t results of every image block
(the ranks meet here once per step), followed by a 100 KB all-reduce of the per-shard "images ahead" counts.
The image blocks are sized by the measured speed of every GPU (``ShardBalancer``): under the 1 kW power cap
the B200s of one box differ by several percent in sustained clock, and equal blocks leave the fast ones
waiting in the collective."""
import collections

import numpy as np
import torch

from . import _cabi, ranking, scoring
from .tiling import build_region_tiles


def shard_bounds(n, world, rank):
    """Contiguous image block of `rank` for equal blocks: [lo, hi)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class ShardBalancer:
    """Image-block bounds proportional to the measured speed of every rank.

    Every step each rank times its own scoring launches (CUDA events) and ships (images scored, milliseconds) of its
    PREVIOUS step inside the ranking exchange; all ranks apply the same update to the same gathered numbers, so the
    bounds stay identical everywhere without an extra collective.  Results do not depend on the partition."""
    SMOOTH = 0.5          # weight of a new measurement
    CLAMP = (0.75, 1.3)   # relative speed is kept inside this band

    def __init__(self):
        self.speed = {}            # world -> float64 [world], mean 1
        self.enabled = True        # False: keep the current bounds (A/B runs, tests)
        self.last = None           # the last gathered (images, ms) lists

    def all_bounds(self, n, world):
        sp = self.speed.get(world)
        if sp is None or n < 8 * world:
            return [shard_bounds(n, world, r) for r in range(world)]
        cum = np.cumsum(sp) / float(np.sum(sp))
        edges = [0] + [int(round(n * c)) for c in cum[:-1]] + [n]
        edges = np.maximum.accumulate(np.clip(edges, 0, n))
        return [(int(edges[r]), int(edges[r + 1])) for r in range(world)]

    def bounds(self, n, world, rank):
        return self.all_bounds(n, world)[rank]

    def update(self, world, images, ms):
        """images / ms: per-rank work and time of one earlier step (entries with ms <= 0 carry no measurement)."""
        images, ms = np.asarray(images, np.float64), np.asarray(ms, np.float64)
        self.last = (images.tolist(), ms.tolist())
        if len(images) != world or np.any(ms <= 0) or np.any(images <= 0) or not self.enabled:
            return
        rate = images / ms
        rate = rate / rate.mean()
        old = self.speed.get(world)
        new = rate if old is None else (1 - self.SMOOTH) * old + self.SMOOTH * rate
        new = np.clip(new, *self.CLAMP)
        self.speed[world] = new / new.mean()

    def reset(self):
        self.speed.clear()


balancer = ShardBalancer()
# Device-resident captions: the HBM-bound pack kernel (1.3 ms for 25 000 captions) crawls next to the persistent tcgen05
# kernel (~12 ms co-scheduled, profiles/r02_n2_phase_balance_ab.md), so a two-phase split with a large second phase
# stalls the scoring (measured at N = 2: 173-177 ms against 166 ms).  With RAMPED phases (Nc/32, 3/32, 9/32, the rest) every
# pack still finishes inside the previous phase's scoring even at the crawling rate, and only the first, 1/32 pack is
# exposed.  Measured at N = 8 (profiles/r02_n8_device_phases_ab.md): 45.2 ms per step with the four launches against
# 43.6 ms with one pack + one launch -- the first launch still starts 1.7-2.0 ms into the step (host enqueue, not the pack,
# sets that), and the later packs finish only just inside the previous launch.  Off; True enables it for A/B runs.
DEVICE_PHASES = False
# this rank's recent scoring passes, oldest first: (images scored, [(start_event, end_event), ...]).  The exchange of
# step k ships the newest pass whose events have COMPLETED (normally step k-1: the launches of step k are still running
# when its payload is assembled) -- never waits, never reads an unfinished event.
_timings = collections.deque(maxlen=4)
# diagnostics (tools/e2e_timeline.py): when a list, the host-resident sharded path appends per caption phase
# dict(c0, c1, t0, packed, gathered [prep-stream events], s0, s1 [main-stream events around the scoring launch])
phase_timeline = None
# diagnostics: when a list, rank_device appends one dict of CUDA events (stage name -> event recorded after that stage)
rank_timeline = None


def _take_timing():
    """(images, ms) of this rank's latest completed scoring pass, (0, 0) when there is none; consumes it and
    everything older."""
    best = None
    while _timings and _timings[0][1][-1][1].query():
        best = _timings.popleft()
    if best is None:
        return 0.0, 0.0
    n, ev = best
    return float(n), float(sum(a.elapsed_time(b) for a, b in ev))


# Cross-GPU work pool of the device-resident sharded pass (steal.py): the last POOL_TAIL of every rank's word units is cut
# into POOL_CHUNKS chunks that any rank may score into the owner's block over NVLink.  False: static blocks only.
# Chunk sizes decrease linearly (12 : 11 : ... : 1 of the tail): a rank keeps two launches queued, so the residual
# imbalance is about two of the LAST chunks (0.2 % of a rank's work each), while the early chunks stay large.
POOL = True
POOL_TAIL = 0.16
POOL_CHUNKS = 12
_pools = {}


def pool_cuts(n_units):
    """Word-unit boundaries of the tail chunks: [u_main, ..., n_units], sizes decreasing linearly, every chunk >= 1 unit."""
    tail = int(round(n_units * POOL_TAIL))
    C = max(1, min(POOL_CHUNKS, tail))
    if n_units < 16 or tail < 1:
        return [0, n_units]                      # too small to split: one chunk = everything
    w = np.arange(C, 0, -1, dtype=np.float64)
    sizes = np.maximum(1, np.floor(w / w.sum() * tail)).astype(np.int64)
    sizes[0] += tail - int(sizes.sum())           # rounding goes to the first (largest) chunk
    if sizes[0] < 1:
        return [n_units - C] + [n_units - C + k + 1 for k in range(C)]
    cuts = [n_units - tail]
    for sz in sizes:
        cuts.append(cuts[-1] + int(sz))
    return cuts


def _work_pool(group, rows, Nc):
    from . import steal
    key = id(group)
    pool = _pools.get(key)
    if pool is not None and not pool.fits(rows, Nc):
        pool.close()
        del _pools[key]
        pool = None
    if pool is None:
        try:
            pool = _pools[key] = steal.WorkPool(group, rows, Nc)
        except _cabi.AladError as e:             # raised on every rank alike (peer.PeerWindow / steal.Counters)
            global POOL
            import warnings
            warnings.warn(f"cross-GPU work pool unavailable ({e}); static image blocks")
            POOL = False
            return None
    return pool


# How the packed caption rows of a phase reach all ranks when the captions live on the host (every rank uploads and
# packs 1/world of them): "peer" = copy engines through IPC peer windows (peer.py; needs NVLink / P2P access between
# the ranks' GPUs, one node), "nccl" = all_gather_into_tensor.
EXCHANGE = "peer"
_exchanges = {}


class _CaptionExchange:
    """Per process group: the peer window (2 x words, 2 x caption ids, ready / ack flags), the side streams and the
    global phase counter of the host-resident sharded path.  Persistent across calls: sequence numbers keep growing,
    so the acknowledgements of one call's last phases gate the first phases of the next."""
    FLAG_BYTES = 4096

    def __init__(self, group, world, pad_max, Kp):
        from . import peer
        self.world, self.pad_max, self.Kp = world, pad_max, Kp
        up = lambda x: (x + 1023) // 1024 * 1024       # noqa: E731
        wbytes, cbytes = up(world * pad_max * Kp * 2), up(world * pad_max * 4)
        self.off_words = [0, wbytes]
        self.off_caps = [2 * wbytes, 2 * wbytes + cbytes]
        self.off_flags = 2 * wbytes + 2 * cbytes
        assert 4 * (2 * 2 * world + 1) <= self.FLAG_BYTES
        self.win = peer.PeerWindow(self.off_flags + self.FLAG_BYTES, group)
        self.words = [self.win.view(o, world * pad_max * Kp * 2, torch.bfloat16).view(world * pad_max, Kp) for o in self.off_words]
        self.caps = [self.win.view(o, world * pad_max * 4, torch.int32) for o in self.off_caps]
        self.error_ptr = self.win.local + self.off_flags + 4 * (4 * world)
        self.error_view = self.win.view(self.off_flags + 4 * (4 * world), 4, torch.int32)
        self.error_host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        self.error_event = None
        self.prep, self.xchg = torch.cuda.Stream(), torch.cuda.Stream()
        self.freed, self.sent = [None, None], [None, None]
        self.g = 0

    def off_flag(self, kind, b, q):
        """Byte offset of flag slot (kind, buffer b, source rank q) inside a window."""
        return self.off_flags + 4 * (((0 if kind == "ready" else 2) + b) * self.world + q)

    def flag_ptr(self, kind, b):
        return self.win.local + self.off_flag(kind, b, 0)

    def check_error_async(self, stream):
        """A wait that timed out (lost peer) leaves 1 + rank in the error slot: raise at the next call."""
        if self.error_event is not None and self.error_event.query() and int(self.error_host[0]) != 0:
            raise _cabi.AladError(f"peer exchange: timed out waiting for rank {int(self.error_host[0]) - 1}")
        self.error_host.copy_(self.error_view, non_blocking=True)
        self.error_event = torch.cuda.Event()
        self.error_event.record(stream)

    def close(self):
        self.win.close()


def _caption_exchange(group, world, pad_max, Kp):
    """The exchange of `group`, (re)built collectively when a call needs larger buffers.  None when peer windows are
    not available (the caller falls back to NCCL)."""
    global EXCHANGE
    key = id(group)
    xc = _exchanges.get(key)
    if xc is not None and (xc.pad_max < pad_max or xc.Kp != Kp or xc.world != world):
        xc.close()
        xc = None
        del _exchanges[key]
    if xc is None:
        try:
            xc = _exchanges[key] = _CaptionExchange(group, world, pad_max, Kp)
        except _cabi.AladError as e:
            import warnings
            warnings.warn(f"peer windows unavailable ({e}); the packed captions are exchanged with NCCL all-gathers")
            EXCHANGE = "nccl"
            return None
    return xc


def close_exchanges():
    """Free the peer windows (collective; call before destroying the process group)."""
    for xc in list(_exchanges.values()):
        xc.close()
    _exchanges.clear()
    for pool in list(_pools.values()):
        pool.close()
    _pools.clear()


# Derived host arrays of a gallery (valid counts, clamp flags) are a function of the python length lists the
# reference passes around (25 000 entries at COCO-5k: ~1 ms of list -> numpy conversion per call); memoised by content.
_META = collections.OrderedDict()


_LENS = []          # most recent first: (copy of img_lens, copy of cap_lens, key)


def lens_key(img_lens, cap_lens):
    """Content key of the two python length lists, computed once per call and shared by the score-block cache of
    evaluation.py and the gallery metadata memo.  Hashing two 25 000-entry lists costs 0.33 ms; comparing them with the
    copies kept from the last calls costs 0.055 ms and is just as exact, so repeated calls on equal lists (i2t then t2i,
    every step of a benchmark) take the comparison."""
    if isinstance(img_lens, np.ndarray):
        img_lens = img_lens.tolist()
    if isinstance(cap_lens, np.ndarray):
        cap_lens = cap_lens.tolist()
    for e in _LENS:
        if len(e[0]) == len(img_lens) and len(e[1]) == len(cap_lens) and e[0] == img_lens and e[1] == cap_lens:
            return e[2]
    key = (len(img_lens), hash(tuple(img_lens)), len(cap_lens), hash(tuple(cap_lens)))
    _LENS.insert(0, (list(img_lens), list(cap_lens), key))
    del _LENS[4:]
    return key


def _gallery_meta(img_shape1, cap_shape, img_lens, cap_lens, Ni, img_start, img_step, lkey=None):
    if isinstance(img_lens, np.ndarray):
        img_lens = img_lens.tolist()
    if isinstance(cap_lens, np.ndarray):
        cap_lens = cap_lens.tolist()
    key = (lkey or lens_key(img_lens, cap_lens)) + (img_shape1, tuple(cap_shape), Ni, img_start, img_step)
    hit = _META.get(key)
    if hit is not None:
        _META.move_to_end(key)
        return hit
    img_lens_g = img_lens[img_start:img_start + (Ni - 1) * img_step + 1:img_step] if Ni else []
    if len(img_lens_g) != Ni:
        raise ValueError("length lists do not match the batch sizes")
    val = scoring.scored_counts((Ni, img_shape1), cap_shape, img_lens_g, cap_lens)
    _META[key] = val
    if len(_META) > 8:
        _META.popitem(last=False)
    return val


STAGING_THREADS_MAX = 16


def staging_threads():
    """Host threads that gather a pageable source into the pinned staging buffers: the CPUs this process may use,
    shared between the ranks of the box, at most STAGING_THREADS_MAX (measured on a 16-core B200 box, profiles/r02_pageable_upload.md:
    9.5 GB/s with 1 thread, 38 GB/s with 8, 44.5 GB/s with 16)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    forced = os.environ.get("ALAD_H2D_THREADS")
    if forced:
        return max(1, int(forced))
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(STAGING_THREADS_MAX, n // max(local, 1)))


def _upload_rows(x, row_start, row_step, n_rows, n_slots, out=None):
    """Rows row_start + i*row_step (i < n_rows), slots [0, n_slots) of a [N,S,d] fp32 tensor ->
    device tensor [n_rows, n_slots, d].  CPU sources go through one pitched H2D copy."""
    N, S, d = x.shape
    if x.is_cuda:
        return x[row_start:row_start + (n_rows - 1) * row_step + 1:row_step, :n_slots] if n_rows else x[:0, :n_slots]
    if x.dtype != torch.float32 or x.stride(2) != 1 or x.stride(1) != d:
        x = x.float().contiguous()
    if out is None:
        out = torch.empty((n_rows, n_slots, d), dtype=torch.float32, device="cuda")
    dst = out[:n_rows, :n_slots]
    assert out.shape[1] == n_slots and out.shape[2] == d and out.is_contiguous()
    if n_rows and n_slots:
        src = x.data_ptr() + row_start * x.stride(0) * 4
        if x.is_pinned():
            _cabi.check(_cabi.lib().alad_h2d_2d(out.data_ptr(), n_slots * d * 4, src, row_step * x.stride(0) * 4,
                                                n_slots * d * 4, n_rows, _cabi.stream_ptr()), "alad_h2d_2d")
        else:
            # pageable source (what the reference's encode_data returns): multi-threaded staging through pinned buffers
            _cabi.check(_cabi.lib().alad_h2d_2d_staged(out.data_ptr(), n_slots * d * 4, src, row_step * x.stride(0) * 4,
                                                       n_slots * d * 4, n_rows, staging_threads(), _cabi.stream_ptr()),
                        "alad_h2d_2d_staged")
    return dst


class AlignmentGallery:
    """Scores a block of gallery images against all captions.

    images   [N_img_rows, S_im, d]  fp32, CPU (pinned or pageable) or CUDA
    captions [Nc, S_s, d]
    image i of the gallery is row img_start + i*img_step (the reference stores every image 5x:
    i2t reads row 5i, t2i rows 0::5 -- alad/evaluation.py:178,252)."""

    def __init__(self, images, captions, img_lens, cap_lens, n_images, img_start=0, img_step=1,
                 precision=None, world=1, rank=0, caption_chunk=4096, caption_phases=None, bounds=None, lkey=None):
        if not torch.cuda.is_available():
            raise _cabi.AladError("aladin_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        from .gallery import DeviceContainer
        self.prepacked = isinstance(images, DeviceContainer)
        if self.prepacked != isinstance(captions, DeviceContainer):
            raise TypeError("images and captions must both be DeviceContainers (gallery.encode_data) or both tensors")
        if self.prepacked:
            if images.precision != captions.precision:
                raise ValueError("image and caption containers were packed with different precisions")
            if (img_start, img_step) not in ((0, 5), (0, 1)):
                raise ValueError("DeviceContainer images are the distinct gallery images (rows 0::5)")
            precision = images.precision            # the packed operands fix the mode
            img_step = 5
        self.precision = precision or scoring.get_precision()
        self.images, self.captions = images, captions
        self.Ni, self.Nc = int(n_images), int(captions.shape[0])
        self.img_start, self.img_step = img_start, img_step
        self.world, self.rank = world, rank
        # image blocks of all ranks: equal blocks for a single rank, speed-weighted ones otherwise (ShardBalancer)
        # the work pool evens out the ranks dynamically: equal blocks; otherwise speed-weighted ones (ShardBalancer)
        pooled = POOL and world > 1 and bounds is None and not self.prepacked and images.is_cuda and captions.is_cuda
        self.equal_blocks = pooled
        if pooled:
            bounds = [shard_bounds(self.Ni, world, r) for r in range(world)]
        self.bounds = bounds if bounds is not None else balancer.all_bounds(self.Ni, world)
        self.lo, self.hi = self.bounds[rank]
        self.caption_chunk = caption_chunk
        self.caption_phases = caption_phases          # None: ramped phases (see _phase_bounds)
        self.R, self.W, self.nr, self.nw, self.clamp = _gallery_meta(
            images.shape[1], tuple(captions.shape), img_lens, cap_lens, self.Ni, img_start, img_step, lkey)
        self._events = []

    def _score(self, words, regions, tiles_dev, n_tiles, n_loc, n_caps, out):
        """One scoring launch, bracketed by events for the shard balancer."""
        if self.world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        scoring.mrsw_scores_packed(words, regions, tiles_dev, n_tiles, n_loc, n_caps, out=out)
        if self.world > 1:
            e1.record()
            self._events.append((e0, e1))

    def _done(self, n_loc):
        if self.world > 1:
            if n_loc and self._events:
                _timings.append((n_loc, self._events))
            self._events = []

    def _phase_bounds(self):
        """Caption-column phases of the host-resident multi-rank path: the upload + pack + exchange of phase p+1
        hides behind the scoring of phase p, so only phase 0 is exposed -- it is small (Nc/64) and the phases grow
        by 1.5x up to Nc/(2*world).  The cap follows the world size because the PCIe upload of a rank's share runs at a
        fixed rate while the scoring time shrinks with 1/world: at 8 ranks the two take about as long (30 ms against 36 ms at
        COCO-5k, profiles/r02_e2e_timeline.md), the pipeline is upload-paced, and whatever is scored after the last
        upload has landed is exposed -- one Nc/16 phase instead of the Nc/3 one of a fixed Nc/4 cap."""
        Nc = self.Nc
        if self.caption_phases is not None:
            P = max(1, self.caption_phases if Nc >= self.caption_phases * self.world * 64 else 1)
            return [(p * Nc // P, (p + 1) * Nc // P) for p in range(P)]
        if Nc < 64 * self.world * 16:
            return [(0, Nc)]
        cap = max(Nc // (2 * max(self.world, 2)), 64 * self.world)
        out, c0, size = [], 0, max(Nc // 64, 64 * self.world)
        while c0 < Nc:
            c1 = min(Nc, c0 + int(size))
            if Nc - c1 < size // 2:
                c1 = Nc
            out.append((c0, c1))
            c0 = c1
            size = min(size * 1.5, cap)
        return out

    def _row_csum(self):
        """First packed word row of every caption in the canonical (unsharded, unphased) packing."""
        if getattr(self, "_csum", None) is None:
            self._csum = np.concatenate([[0], np.cumsum(self.nw, dtype=np.int64)])
        return self._csum

    def _phase_plans(self):
        """(phase bounds, per phase (this rank's caption span, padded rows per rank slot, first row inside the slot),
        largest slot) of the host-resident sharded path: inside a phase every rank uploads + packs its 1/world share.
        A share starts at row (canonical row of its first caption) mod 256 of its slot, so that the 256-row work units
        of the scoring kernel cut every caption exactly where the unsharded single launch cuts it: the <= 2 partial sums
        of an S entry are the same numbers and the scores are bit-identical for any world size and phase split."""
        unit = 2 * _cabi.TILE_M
        csum = self._row_csum()
        pb = self._phase_bounds()
        plans = []
        for c0, c1 in pb:
            spans = [tuple(c0 + x for x in shard_bounds(c1 - c0, self.world, r)) for r in range(self.world)]
            need = max(int(csum[a] % unit + csum[b] - csum[a]) for a, b in spans)
            pad = max(((need + unit - 1) // unit) * unit, unit)
            a = spans[self.rank][0]
            plans.append((spans[self.rank], pad, int(csum[a] % unit)))
        return pb, plans, max(pad for _, pad, _ in plans)

    def _pack_caption_range(self, c_lo, c_hi, words_buf, cap_buf, row_base, split, dev, item_origin=0):
        """Upload (if on the host) and pack captions [c_lo, c_hi) into rows row_base.. of
        words_buf / cap_buf.  Host sources are double-buffered: the pitched H2D copy of chunk k+1
        overlaps the packing of chunk k.  Returns the number of rows written."""
        nw = self.nw
        on_cpu = not self.captions.is_cuda
        chunk = self.caption_chunk if on_cpu else max(c_hi - c_lo, 1)
        bounds = [(c0, min(c_hi, c0 + chunk)) for c0 in range(c_lo, c_hi, chunk)]
        main = torch.cuda.current_stream()
        if on_cpu and bounds:
            Lw_max = 1 + int(nw[c_lo:c_hi].max())
            stage = [torch.empty((chunk, Lw_max, self.captions.shape[2]), dtype=torch.float32, device=dev) for _ in range(2)]
            copy_stream = torch.cuda.Stream()
            copy_stream.wait_stream(main)
            ready = [torch.cuda.Event() for _ in bounds]
            freed = [None, None]
        rows = 0
        for k, (c0, c1) in enumerate(bounds):
            if on_cpu:
                bsel = k & 1
                with torch.cuda.stream(copy_stream):
                    if freed[bsel] is not None:
                        copy_stream.wait_event(freed[bsel])
                    cap_dev = _upload_rows(self.captions, c0, 1, c1 - c0, Lw_max, out=stage[bsel])
                    ready[k].record(copy_stream)
                main.wait_event(ready[k])
            else:
                cap_dev = self.captions[c0:c1]
            scoring.pack_tokens(cap_dev, nw[c0:c1], slot0=1, mode=scoring.WORD_MODE[split], out=words_buf, out_row_item=cap_buf,
                                row_base=row_base + rows, item_base=c0 - item_origin)
            rows += int(nw[c0:c1].sum())
            if on_cpu:
                freed[bsel] = torch.cuda.Event()
                freed[bsel].record(main)
        return rows

    def _scores_pooled(self, group, split, dev):
        """Device-resident sharded pass under the cross-GPU work pool (steal.py): every rank packs ALL regions (0.35 GB
        at COCO-5k, 0.2 ms) and all captions, scores the main part of its own block with one launch and the tail in
        chunks; idle ranks take chunks of the rank with the most left and add them into its block over NVLink."""
        from . import steal
        from .tiling import exclusive_cumsum
        W, r = self.world, self.rank
        Nc, nw = self.Nc, self.nw
        rows_max = max(b - a for a, b in self.bounds)
        pool = _work_pool(group, rows_max, Nc)
        if pool is None:                        # no peer access between the ranks: the static path (equal blocks)
            return self.scores(group=None)
        n_loc = self.hi - self.lo
        blk = pool.block(n_loc)
        # ---- operands: all regions, all words
        Lr = 1 + int(self.nr.max()) if self.Ni else 1
        im_dev = _upload_rows(self.images, self.img_start, self.img_step, self.Ni, Lr)
        regions = scoring.pack_tokens(im_dev, self.nr, slot0=1, mode=scoring.REGION_MODE[split])
        roff, _ = exclusive_cumsum(self.nr)
        roff = np.concatenate([roff, [regions.n_rows]])
        tables, spans = [], []
        for lo, hi in self.bounds:
            _, table, _ = build_region_tiles(self.nr[lo:hi], self.clamp[lo:hi])
            spans.append((sum(len(t) for t in tables), len(table)))
            tables.append(table)
        flat = np.concatenate([t.reshape(-1) for t in tables]) if sum(len(t) for t in tables) else np.zeros(1, np.uint32)
        tiles_all = scoring._to_dev(flat.view(np.int32), dev)
        words = scoring.pack_tokens(self.captions, nw, slot0=1, mode=scoring.WORD_MODE[split], want_row_item=True)
        n_rows, Kp = words.n_rows, words.Kp
        unit = 2 * _cabi.TILE_M
        n_units = (n_rows + unit - 1) // unit
        cuts = pool_cuts(n_units)
        C, u_main = len(cuts) - 1, cuts[0]

        def launch(owner, u0, u1):
            lo, hi = self.bounds[owner]
            t0, nt = spans[owner]
            r0, r1 = u0 * unit, min(u1 * unit, n_rows)
            if hi <= lo or nt == 0 or r1 <= r0:
                return
            w = scoring.Packed(words.data[r0:r1], r1 - r0, Kp, None, None, words.row_item[r0:u1 * unit], words.mode)
            g = scoring.Packed(regions.data[int(roff[lo]):int(roff[hi])], int(roff[hi] - roff[lo]), Kp, None, None, None, regions.mode)
            scoring.mrsw_scores_packed(w, g, tiles_all[t0 * _cabi.NTILE_WORDS:], nt, hi - lo, Nc, accumulate=True,
                                       out_ptr=pool.win.ptrs[owner], timeline_nc=Nc * (r1 - r0) / max(n_rows, 1))

        def zero_event():
            blk.zero_()
            return steal._cuda_event()

        steal.run(pool, C, lambda: launch(r, 0, u_main), lambda owner, c: launch(owner, cuts[c], cuts[c + 1]), n_loc, zero_event)
        return blk.clone()

    def _score_phases_peer(self, xc, pb, plans, prep_regions, n_loc, S, split, dev):
        """Phase loop with the packed rows replicated by COPY ENGINES through peer windows (peer.py): three streams per
        rank -- `prep` uploads + packs this rank's share of phase g straight into its slot of the local window,
        `xchg` pushes that slot into every peer's window and raises their ready flags, the main stream waits for the
        flags and scores.  Nothing here needs an SM while the scoring kernel of phase g-1 runs, so the exchange of a
        phase hides behind the previous phase's scoring; buffers alternate (g & 1) and a slot is overwritten only after
        every peer has acknowledged that it scored the phase that used it.  The image block (`prep_regions`: upload +
        pack on the main stream) is enqueued after the first phase's caption upload, so that phase's exchange runs
        while the images cross PCIe."""
        from . import peer
        regions = tiles_dev = None
        n_tiles, first = 0, True
        W, r, Kp = self.world, self.rank, xc.Kp
        main = torch.cuda.current_stream()
        prep, xchg = xc.prep, xc.xchg
        prep.wait_stream(main)
        others = [q for q in range(W) if q != r]
        for (c0, c1), ((m0, m1), pad, off) in zip(pb, plans):
            g = xc.g
            xc.g += 1
            b, seq = g & 1, g + 1
            tl = None
            if phase_timeline is not None:
                tl = dict(c0=c0, c1=c1, **{k: torch.cuda.Event(enable_timing=True) for k in ("t0", "packed", "gathered", "s0", "s1")})
                phase_timeline.append(tl)
            words_b, caps_b = xc.words[b], xc.caps[b]
            rows_mine = int(self.nw[m0:m1].sum())
            with torch.cuda.stream(prep):
                if xc.freed[b] is not None:
                    prep.wait_event(xc.freed[b])          # local scoring of phase g-2 has read this buffer
                if xc.sent[b] is not None:
                    prep.wait_event(xc.sent[b])           # ... and its slot has left for the peers
                if tl:
                    tl["t0"].record(prep)
                mw, mc = words_b[r * pad:(r + 1) * pad], caps_b[r * pad:(r + 1) * pad]
                mc.fill_(-1)                              # gap rows are never scored (row_cap -1)
                self._pack_caption_range(m0, m1, mw, mc, off, split, dev, item_origin=c0)
                packed = torch.cuda.Event(enable_timing=tl is not None)
                packed.record(prep)
                if tl:
                    tl["packed"] = packed
            with torch.cuda.stream(xchg):
                xchg.wait_event(packed)
                if g >= 2:                                # every peer has scored phase g-2 out of its buffer b
                    peer.wait(xc.flag_ptr("ack", b), W, seq - 2, r, xc.error_ptr, xchg)
                w_off, c_off = xc.off_words[b] + (r * pad + off) * Kp * 2, xc.off_caps[b] + r * pad * 4
                for q in others:
                    peer.copy(xc.win.ptrs[q] + w_off, xc.win.local + w_off, rows_mine * Kp * 2, xchg)
                    peer.copy(xc.win.ptrs[q] + c_off, xc.win.local + c_off, pad * 4, xchg)
                peer.signal([xc.win.ptrs[q] + xc.off_flag("ready", b, r) for q in others], seq, xchg)
                xc.sent[b] = torch.cuda.Event()
                xc.sent[b].record(xchg)
            if first:
                regions, tiles_dev, n_tiles = prep_regions()
                first = False
            main.wait_event(packed)
            peer.wait(xc.flag_ptr("ready", b), W, seq, r, xc.error_ptr, main)
            if tl:
                tl["gathered"].record(main)
                tl["s0"].record(main)
            if n_loc:
                words = scoring.Packed(words_b[:W * pad], W * pad, Kp, None, None, caps_b[:W * pad], scoring.WORD_MODE[split])
                self._score(words, regions, tiles_dev, n_tiles, n_loc, c1 - c0, S[:, c0:c1])
            if tl:
                tl["s1"].record(main)
            xc.freed[b] = torch.cuda.Event()
            xc.freed[b].record(main)
            peer.signal([xc.win.ptrs[q] + xc.off_flag("ack", b, r) for q in others], seq, main)
        xc.check_error_async(main)

    def _score_phases_nccl(self, group, pb, plans, pad_max, Kp, regions, tiles_dev, n_tiles, n_loc, S, split, dev):
        """Phase loop with an NCCL all-gather of the packed rows.  The all-gather kernel cannot start next to the
        persistent scoring kernel, so every phase's exchange lands in the gap after the previous phase's scoring
        (profiles/r02_e2e_timeline.md); kept for process groups without peer access (EXCHANGE = 'nccl')."""
        import torch.distributed as dist
        words_buf = [torch.empty((self.world * pad_max, Kp), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        caps_buf = [torch.empty((self.world * pad_max,), dtype=torch.int32, device=dev) for _ in range(2)]
        mine_w = [torch.empty((pad_max, Kp), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        mine_c = [torch.empty((pad_max,), dtype=torch.int32, device=dev) for _ in range(2)]
        main = torch.cuda.current_stream()
        prep = torch.cuda.Stream()
        prep.wait_stream(main)
        freed = [None, None]
        for p, ((c0, c1), ((m0, m1), pad, off)) in enumerate(zip(pb, plans)):
            b = p & 1
            tl = None
            if phase_timeline is not None:
                tl = dict(c0=c0, c1=c1, **{k: torch.cuda.Event(enable_timing=True) for k in ("t0", "packed", "gathered", "s0", "s1")})
                phase_timeline.append(tl)
            with torch.cuda.stream(prep):
                if freed[b] is not None:
                    prep.wait_event(freed[b])
                if tl:
                    tl["t0"].record(prep)
                mw, mc = mine_w[b][:pad], mine_c[b][:pad]
                mc.fill_(-1)
                self._pack_caption_range(m0, m1, mw, mc, off, split, dev, item_origin=c0)
                if tl:
                    tl["packed"].record(prep)
                wa, ca = words_buf[b][:self.world * pad], caps_buf[b][:self.world * pad]
                dist.all_gather_into_tensor(wa, mw, group=group)      # gap rows are never scored (row_cap -1)
                dist.all_gather_into_tensor(ca, mc, group=group)
                ready = torch.cuda.Event(enable_timing=tl is not None)
                ready.record(prep)
                if tl:
                    tl["gathered"] = ready
            main.wait_event(ready)
            if tl:
                tl["s0"].record(main)
            if n_loc:
                words = scoring.Packed(wa, self.world * pad, Kp, None, None, ca, scoring.WORD_MODE[split])
                self._score(words, regions, tiles_dev, n_tiles, n_loc, c1 - c0, S[:, c0:c1])
            if tl:
                tl["s1"].record(main)
            freed[b] = torch.cuda.Event()
            freed[b].record(main)

    def packed_operands(self):
        """Device-resident packed operands of this shard: (words Packed over ALL captions incl. row_item,
        regions Packed over the image block, first packed region row of every local image [int64 numpy]).
        Used by the pair-list (two-stage) path, which scores many small tiles from the same operands."""
        from .tiling import exclusive_cumsum, padded_rows
        split = scoring.PRECISION_CODE[self.precision]      # 0 bf16, 1 split-precision fp32, 2 tf32
        lo, hi = self.lo, self.hi
        n_loc = hi - lo
        dev = torch.device("cuda", torch.cuda.current_device())
        nr = self.nr[lo:hi]
        row_off, _ = exclusive_cumsum(nr)
        if self.prepacked:
            regions, nr_loc = self.images.rows_of(lo, hi)
            assert np.array_equal(nr_loc, nr)
            return self.captions.packed, regions, row_off
        d = self.captions.shape[2]
        Kp = ((d * scoring.K_FACTOR[split] + _cabi.TILE_K - 1) // _cabi.TILE_K) * _cabi.TILE_K
        regions = None
        if n_loc:
            Lr = 1 + int(nr.max())
            im_dev = _upload_rows(self.images, self.img_start + lo * self.img_step, self.img_step, n_loc, Lr)
            regions = scoring.pack_tokens(im_dev, nr, slot0=1, mode=scoring.REGION_MODE[split])
        n_rows = int(self.nw.sum())
        words_buf = torch.empty((max(n_rows, 1), Kp), dtype=torch.bfloat16, device=dev)
        cap_buf = torch.full((max(padded_rows(n_rows), 2 * _cabi.TILE_M),), -1, dtype=torch.int32, device=dev)
        if self.Nc:
            self._pack_caption_range(0, self.Nc, words_buf, cap_buf, 0, split, dev)
        words = scoring.Packed(words_buf, n_rows, Kp, self.nw, None, cap_buf, scoring.WORD_MODE[split])
        return words, regions, row_off

    def scores(self, group=None):
        """S[hi-lo, Nc] fp32 on the device for this shard's image block.

        With a process group (world > 1) and captions on the HOST, every rank uploads and packs
        only its 1/world share of the captions and the packed bf16 rows are all-gathered over
        NVLink (2.6 GB in total at COCO-5k) instead of every rank pulling all 5 GB over PCIe."""
        import torch.distributed as dist
        split = scoring.PRECISION_CODE[self.precision]      # 0 bf16, 1 split-precision fp32, 2 tf32
        lo, hi = self.lo, self.hi
        n_loc = hi - lo
        dev = torch.device("cuda", torch.cuda.current_device())
        S = torch.empty((n_loc, self.Nc), dtype=torch.float32, device=dev)
        if self.prepacked:
            # backbone outputs were packed on the device by gallery.GalleryWriter: no upload, no pack
            if n_loc == 0 or self.Nc == 0:
                return S
            regions, nr_loc = self.images.rows_of(lo, hi)
            assert np.array_equal(nr_loc, self.nr[lo:hi]) and np.array_equal(self.captions.counts, self.nw)
            _, table, _ = build_region_tiles(nr_loc, self.clamp[lo:hi])
            tiles_dev = scoring._to_dev(table.view(np.int32).reshape(-1), dev) if len(table) else None
            self._score(self.captions.packed, regions, tiles_dev, len(table), n_loc, self.Nc, S)
            self._done(n_loc)
            return S
        shard_caps = (group is not None and self.world > 1 and not self.captions.is_cuda and dist.is_initialized())
        if (POOL and group is not None and self.world > 1 and dist.is_initialized() and self.captions.is_cuda
                and self.images.is_cuda and self.Nc > 0 and self.equal_blocks):
            return self._scores_pooled(group, split, dev)
        if (n_loc == 0 and not shard_caps) or self.Nc == 0:
            return S
        nr, nw = self.nr[lo:hi], self.nw
        d = self.captions.shape[2]
        Kp = ((d * scoring.K_FACTOR[split] + _cabi.TILE_K - 1) // _cabi.TILE_K) * _cabi.TILE_K
        # ---- regions of this image block
        def prep_regions():
            if not n_loc:
                return None, None, 0
            Lr = 1 + int(nr.max())
            im_dev = _upload_rows(self.images, self.img_start + lo * self.img_step, self.img_step, n_loc, Lr)
            reg = scoring.pack_tokens(im_dev, nr, slot0=1, mode=scoring.REGION_MODE[split])
            _, table, _ = build_region_tiles(nr, self.clamp[lo:hi])
            return reg, (scoring._to_dev(table.view(np.int32).reshape(-1), dev) if len(table) else None), len(table)

        peer_path = False
        if shard_caps and EXCHANGE == "peer":
            # the image block is uploaded AFTER the first caption phase has been enqueued (see _score_phases_peer)
            pb, plans, pad_max = self._phase_plans()
            xc = _caption_exchange(group, self.world, pad_max, Kp)
            peer_path = xc is not None
        if peer_path:
            self._score_phases_peer(xc, pb, plans, prep_regions, n_loc, S, split, dev)
            self._done(n_loc)
            return S
        regions, tiles_dev, n_tiles = prep_regions()
        # ---- words: all captions (single rank / device-resident) or this rank's share + all-gather
        if shard_caps:
            pb, plans, pad_max = self._phase_plans()
            self._score_phases_nccl(group, pb, plans, pad_max, Kp, regions, tiles_dev, n_tiles, n_loc, S, split, dev)
            self._done(n_loc)
            return S
        else:
            # one rank prepares all captions, in column phases: phase k is scored while phase k+1 is uploaded (host
            # sources) and packed on a side stream, so only the first, small phase is exposed
            from .tiling import padded_rows
            on_cpu = not self.captions.is_cuda
            if on_cpu:
                # the first phases are small (chunk/8, /4, /2) so that scoring starts after a few ms of PCIe traffic
                # instead of a whole chunk's worth; the copies stay ahead of the scoring from then on
                chunk = self.caption_chunk
                bounds, c0, size = [], 0, min(chunk, max(chunk // 8, 256))
                while c0 < self.Nc:
                    bounds.append((c0, min(self.Nc, c0 + size)))
                    c0 += size
                    size = min(chunk, 2 * size)
            elif self.Nc >= 4096 and DEVICE_PHASES:
                cuts = [0, self.Nc // 32, self.Nc // 8, 13 * self.Nc // 32, self.Nc]
                bounds = list(zip(cuts[:-1], cuts[1:]))
            else:
                bounds = [(0, self.Nc)]
            if len(bounds) == 1 and not on_cpu:
                words = scoring.pack_tokens(self.captions, nw, slot0=1, mode=scoring.WORD_MODE[split], want_row_item=True)
                self._score(words, regions, tiles_dev, n_tiles, n_loc, self.Nc, S)
                self._done(n_loc)
                return S
            # every phase starts at row (canonical first row) mod 256 of its buffer: same cuts as the single launch (see _phase_plans)
            unit = 2 * _cabi.TILE_M
            csum = self._row_csum()
            rows_max = max(int(csum[c0] % unit + csum[c1] - csum[c0]) for c0, c1 in bounds)
            n_buf = 2 if len(bounds) > 1 else 1
            words_buf = [torch.empty((max(rows_max, 1), Kp), dtype=torch.bfloat16, device=dev) for _ in range(n_buf)]
            cap_buf = [torch.empty((max(padded_rows(rows_max), 2 * _cabi.TILE_M),), dtype=torch.int32, device=dev)
                       for _ in range(n_buf)]
            main = torch.cuda.current_stream()
            prep = torch.cuda.Stream()
            prep.wait_stream(main)
            freed = [None] * n_buf
            for k, (c0, c1) in enumerate(bounds):
                b = k % n_buf
                with torch.cuda.stream(prep):
                    if freed[b] is not None:
                        prep.wait_event(freed[b])
                    cap_buf[b].fill_(-1)
                    off = int(csum[c0] % unit)
                    rows = off + self._pack_caption_range(c0, c1, words_buf[b], cap_buf[b], off, split, dev, item_origin=c0)
                    ready = torch.cuda.Event()
                    ready.record(prep)
                main.wait_event(ready)
                words = scoring.Packed(words_buf[b], rows, Kp, None, None, cap_buf[b], scoring.WORD_MODE[split])
                self._score(words, regions, tiles_dev, n_tiles, n_loc, c1 - c0, S[:, c0:c1])
                freed[b] = torch.cuda.Event()
                freed[b].record(main)
            self._done(n_loc)
            return S


def rank_device(S, npts, img_off=0, n_images_total=None, k=50, group=None, ops=ranking, bounds=None):
    """Exact i2t / t2i ranks and top-k from a shard's score block S[n_loc, Nc], as DEVICE tensors:
    (rank_i2t[npts] int32, top1[npts] int32, rank_t2i[ncq] int32, topk_score[ncq, k], topk_idx[ncq, k] int32,
    timing [world, 2] float32 or None: every rank's (images, ms) of its previous scoring pass).

    `ops` provides rank_rows / col_gt / col_count / col_topk / topk_merge (the CUDA kernels by default; the gloo
    CPU tests plug the oracle in to exercise the exchange protocol).  `bounds` = image blocks [(lo, hi)] of all
    ranks (default: equal blocks)."""
    import torch.distributed as dist
    n_loc, Nc = S.shape
    dist_on = group is not None and dist.is_initialized() and dist.get_world_size(group) > 1
    Ni_total = n_images_total if n_images_total is not None else n_loc
    npts = min(npts, Ni_total)
    marks = {} if rank_timeline is not None and S.is_cuda else None

    def mark(name):
        if marks is not None:
            marks[name] = torch.cuda.Event(enable_timing=True)
            marks[name].record()

    if marks is not None:
        rank_timeline.append(marks)
    mark("start")
    # ---------------- i2t: rows (queries = images < npts), gallery = all captions
    q_loc = max(0, min(n_loc, npts - img_off))
    # ---------------- t2i: columns (queries = captions < 5*npts), gallery = all images
    ncq = min(Nc, 5 * npts)
    Sq = S[:, :ncq]
    fused = getattr(ops, "rank_fused", None)
    if fused is not None and S.is_cuda:
        # one sweep of S for the i2t ranks, the group maxima of the top-k select and (single shard) the t2i counts, one for
        # the top-k candidates -- instead of the four sweeps of rank_rows / col_count / col_topk
        if not dist_on:
            rank_i, top1_i, count, ts, ti = fused(S, k, img_off, q_loc, ncq)
            mark("rank_fused")
            return rank_i, top1_i, count, ts, ti, None
        gt = torch.zeros(ncq, dtype=torch.float32, device=S.device)
        ops.col_gt(Sq, gt, 5, img_off)
        rank_i, top1_i, _, ts, ti = fused(S, k, img_off, q_loc, ncq, count=False)
        mark("rank_rows")
        mark("col_gt_topk")
    else:
        rank_i, top1_i = ops.rank_rows(S[:q_loc], 5, img_off)
        mark("rank_rows")
        gt = torch.zeros(ncq, dtype=torch.float32, device=S.device)
        ops.col_gt(Sq, gt, 5, img_off)
        cs, ci = ops.col_topk(Sq, k, img_off)
        ts, ti = ops.topk_merge(cs, ci)
        mark("col_gt_topk")
        if not dist_on:
            count = ops.col_count(Sq, gt, 5, img_off)
            mark("col_count")
            return rank_i, top1_i, count, ts, ti, None
    world = dist.get_world_size(group)
    if bounds is None:
        bounds = [shard_bounds(Ni_total, world, r) for r in range(world)]
    per = max(hi - lo for lo, hi in bounds)
    # ONE all-gather: ground-truth scores (every entry is owned by exactly one shard, 0 elsewhere), the shard's
    # k best (score, image) per caption, its i2t results, and the (images, ms) of its previous scoring pass
    t_n, t_ms = _take_timing()
    tail = torch.full((2 * per + 2,), -1, dtype=torch.int32, device=S.device)
    tail[:q_loc] = rank_i
    tail[per:per + q_loc] = top1_i
    tail[2 * per:] = torch.tensor([t_n, t_ms], dtype=torch.float32).view(torch.int32).to(S.device, non_blocking=True)
    ts_c, ti_c = ts.contiguous(), ti.contiguous()
    payload = torch.cat([gt.view(torch.int32), ts_c.view(torch.int32).reshape(-1), ti_c.reshape(-1), tail])
    n_pay = payload.numel()
    gathered = torch.empty((world * n_pay,), dtype=torch.int32, device=S.device)
    mark("payload")
    dist.all_gather_into_tensor(gathered, payload, group=group)
    mark("all_gather")
    gathered = gathered.view(world, n_pay)
    o = 0
    gt = gathered[:, o:o + ncq].view(torch.float32).sum(dim=0)
    o += ncq
    gs = gathered[:, o:o + ncq * k].view(torch.float32).reshape(world, ncq, k)
    o += ncq * k
    gi = gathered[:, o:o + ncq * k].reshape(world, ncq, k)
    o += ncq * k
    ts, ti = ops.topk_merge(gs.contiguous(), gi.contiguous())
    g_tail = gathered[:, o:]
    rank_i = torch.cat([g_tail[r, :bounds[r][1] - bounds[r][0]] for r in range(world)])[:npts]
    top1_i = torch.cat([g_tail[r, per:per + bounds[r][1] - bounds[r][0]] for r in range(world)])[:npts]
    # images ahead of the ground truth: local count against the global gt, summed over the shards
    mark("merge")
    count = ops.col_count(Sq, gt, 5, img_off)
    mark("col_count")
    dist.all_reduce(count, group=group)
    mark("all_reduce")
    # the gathered (images, ms) pairs update the shard balancer identically on every rank (read with the results)
    return rank_i, top1_i, count, ts, ti, g_tail[:, 2 * per:2 * per + 2].view(torch.float32)


def rank_both_directions(S, npts, img_off=0, n_images_total=None, k=50, group=None, gather_i2t=True, ops=ranking,
                         bounds=None, lists_to_host=True):
    """rank_device + ONE device->host copy.  Returns numpy arrays shaped like the reference's
    (alad/evaluation.py:166-167,255-256): ranks_i2t[npts], top1[npts], ranks_t2i[5*npts], topk[5*npts, k] (float64).
    lists_to_host=False leaves top1 / topk on the device (int32 tensors; `lists_host` converts them on demand): the
    reference's callers (alad/test.py:271-276, alad/train.py:504-509) only read the metrics, and topk is 10 MB at COCO-5k."""
    rank_i, top1_i, count, _, ti, timing = rank_device(S, npts, img_off, n_images_total, k, group, ops, bounds)
    small = (rank_i, count) if not lists_to_host else (rank_i, top1_i, count, ti)
    if timing is None:
        out = _to_host_f64(*small)
    else:
        out = _to_host_f64(*small, timing.reshape(-1))
        tm = out[-1].reshape(-1, 2)
        balancer.update(tm.shape[0], tm[:, 0], tm[:, 1])
        out = out[:-1]
    if lists_to_host:
        return out
    return out[0], top1_i, out[1], ti


def lists_host(top1_dev, topk_dev):
    """(top1, topk) float64 numpy arrays from the device tensors rank_both_directions(lists_to_host=False) returned."""
    if isinstance(top1_dev, np.ndarray):
        return top1_dev, topk_dev
    return _to_host_f64(top1_dev, topk_dev)


def streaming_ranks(images, captions, img_lens, cap_lens, n_images, img_start=0, img_step=1, precision=None, block_images=1024,
                    k=50, keep_scores=False):
    """Both directions' ranks and top-k WITHOUT materialising S[Ni, Nc] (alad_mrsw_retrieval, one native call): the gallery
    images are scored block by block into one reusable [block_images, Nc] buffer, exactly the way the multi-GPU path scores one
    block per rank -- i2t ranks are final per block, the t2i "images ahead" counts add up over the blocks and the per-caption
    top-k lists are merged as they come.  The ground-truth scores the counts compare against come from a first pass over the
    block diagonal (image block x its own captions, 1 / n_blocks of the work) launched on the same 256-row word units as the
    full pass, so they are bit-identical to the entries of S they stand for.  For galleries whose score matrix does not fit
    (0.5 GB at COCO-5k, 50 GB at ten times the gallery); results equal rank_both_directions on the dense matrix.
    keep_scores=True writes the dense matrix as well (one pass, one ranking) and returns it as a fifth value.
    Returns (ranks_i2t[Ni], top1[Ni], ranks_t2i[Nc], topk[Nc, k]) as float64 numpy arrays."""
    import ctypes as C
    Ni = int(n_images)
    if Ni <= 0:
        raise ValueError("streaming_ranks needs at least one gallery image")
    gal = AlignmentGallery(images, captions, img_lens, cap_lens, n_images=Ni, img_start=img_start, img_step=img_step,
                           precision=precision, world=1, rank=0, bounds=[(0, Ni)])
    Nc = gal.Nc
    k = min(k, Ni)
    words, regions, _ = gal.packed_operands()
    dev = words.data.device
    nr = np.ascontiguousarray(gal.nr, dtype=np.int32)
    clamp = np.ascontiguousarray(gal.clamp, dtype=np.uint8)
    cap_row = np.ascontiguousarray(gal._row_csum(), dtype=np.int64)
    lib = _cabi.lib()
    rank_i = torch.empty(Ni, dtype=torch.int32, device=dev)
    top1_i = torch.empty(Ni, dtype=torch.int32, device=dev)
    count = torch.empty(Nc, dtype=torch.int32, device=dev)
    ts = torch.empty((Nc, k), dtype=torch.float32, device=dev)
    ti = torch.empty((Nc, k), dtype=torch.int32, device=dev)
    S = torch.empty((Ni, Nc), dtype=torch.float32, device=dev) if keep_scores else None
    nbytes = int(lib.alad_mrsw_retrieval_workspace_bytes(Ni, Nc, k, int(block_images), 1 if keep_scores else 0))
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    a = _cabi.MrswRetrievalArgs(
        words=words.data.data_ptr(), n_word_rows=words.n_rows, row_cap=words.row_item.data_ptr(),
        regions=regions.data.data_ptr() if regions is not None else None,
        n_region_rows=regions.n_rows if regions is not None else 0, Kp=words.Kp,
        operand_format=1 if words.mode == 3 else 0, nr=nr.ctypes.data, clamp=clamp.ctypes.data, cap_row=cap_row.ctypes.data,
        Ni=Ni, Nc=Nc, group=5, k=k, block_images=int(block_images),
        S=S.data_ptr() if S is not None else None, ldS=Nc,
        rank_i2t=rank_i.data_ptr(), top1=top1_i.data_ptr(), rank_t2i=count.data_ptr(), topk_score=ts.data_ptr(),
        topk_idx=ti.data_ptr(), workspace=ws.data_ptr(), workspace_bytes=ws.numel())
    _cabi.check(lib.alad_mrsw_retrieval(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_retrieval")
    n_blocks = 1 if keep_scores else -(-Ni // max(int(block_images), k, 1))
    _cabi.launch_count["kernels"] += n_blocks * 9 + (0 if keep_scores else n_blocks * 2)
    out = _to_host_f64(rank_i, top1_i, count, ti)
    return out + (S,) if keep_scores else out


def _to_host_f64(*tensors):
    """Device int tensors -> float64 numpy arrays (the reference returns numpy.zeros-typed arrays:
    alad/evaluation.py:166-167,255-256) with ONE device->host copy into pinned memory and one sync."""
    if not tensors[0].is_cuda:
        return tuple(t.numpy().astype(np.float64) for t in tensors)
    sizes = [t.numel() for t in tensors]
    flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
    host = torch.empty(flat.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(flat, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out, o = [], 0
    arr = host.numpy()
    for t, n in zip(tensors, sizes):
        out.append(arr[o:o + n].reshape(tuple(t.shape)))
        o += n
    return tuple(out)


def _median_of_counts(ranks):
    """numpy.median of an array of non-negative INTEGER values (ranks are counts) from a histogram: 0.07 ms instead
    of the 0.5 ms numpy.partition takes on 25 000 entries.  None when the values are not such integers."""
    n = ranks.size
    if n == 0:
        return None
    ri = ranks.astype(np.int64)
    if ri.min() < 0 or ri.max() > 4 * n + 65536 or not np.array_equal(ri, ranks):
        return None
    cum = np.cumsum(np.bincount(ri))
    lo = int(np.searchsorted(cum, (n - 1) // 2, side="right"))
    hi = int(np.searchsorted(cum, n // 2, side="right"))
    return 0.5 * (lo + hi)


def recall_tuple(ranks):
    """(r1, r5, r10, medr, meanr) exactly as alad/evaluation.py:231-235."""
    n = ranks.size
    r1 = 100.0 * np.count_nonzero(ranks < 1) / n
    r5 = 100.0 * np.count_nonzero(ranks < 5) / n
    r10 = 100.0 * np.count_nonzero(ranks < 10) / n
    med = _median_of_counts(ranks)
    medr = np.floor(np.median(ranks) if med is None else np.float64(med)) + 1
    meanr = ranks.mean() + 1
    return r1, r5, r10, medr, meanr
