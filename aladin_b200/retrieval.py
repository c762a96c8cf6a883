"""All-pairs retrieval engine: upload -> pack -> fused MrSw scores -> exact ranks + top-k.

This is the fused replacement for the per-query Python loops of alad/evaluation.py:175-223
(i2t) and :263-313 (t2i): the gallery goes to the device once, one pass over the
Ni x Nc score block serves both directions, ranks come from "count ahead of the ground
truth" kernels instead of per-query numpy argsorts.

Multi-GPU: gallery images are split into contiguous image blocks, captions are
replicated.  i2t needs no exchange; t2i exchanges the ground-truth scores (all-reduce of
Nc floats), the per-shard counts (all-reduce of Nc ints) and the per-shard top-k candidates
(all-gather of k (score, index) pairs per caption) through torch.distributed / NCCL.
"""
import numpy as np
import torch

from . import _cabi, ranking, scoring
from .tiling import build_region_tiles


def shard_bounds(n, world, rank):
    """Contiguous image block of `rank`: [lo, hi)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


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
        _cabi.check(_cabi.lib().alad_h2d_2d(out.data_ptr(), n_slots * d * 4, src, row_step * x.stride(0) * 4,
                                            n_slots * d * 4, n_rows, _cabi.stream_ptr()), "alad_h2d_2d")
    return dst


class AlignmentGallery:
    """Scores a block of gallery images against all captions.

    images   [N_img_rows, S_im, d]  fp32, CPU (pinned or pageable) or CUDA
    captions [Nc, S_s, d]
    image i of the gallery is row img_start + i*img_step (the reference stores every image 5x:
    i2t reads row 5i, t2i rows 0::5 -- alad/evaluation.py:178,252)."""

    def __init__(self, images, captions, img_lens, cap_lens, n_images, img_start=0, img_step=1,
                 precision=None, world=1, rank=0, caption_chunk=4096, caption_phases=4):
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
        self.lo, self.hi = shard_bounds(self.Ni, world, rank)
        self.caption_chunk = caption_chunk
        self.caption_phases = max(1, caption_phases)
        img_lens_g = [img_lens[img_start + i * img_step] for i in range(self.Ni)]
        self.R, self.W, self.nr, self.nw, self.clamp = scoring.scored_counts(
            (self.Ni, images.shape[1]), captions.shape, img_lens_g, cap_lens)

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
            scoring.pack_tokens(cap_dev, nw[c0:c1], slot0=1, mode=1 if split else 0, out=words_buf, out_row_item=cap_buf,
                                row_base=row_base + rows, item_base=c0 - item_origin)
            rows += int(nw[c0:c1].sum())
            if on_cpu:
                freed[bsel] = torch.cuda.Event()
                freed[bsel].record(main)
        return rows

    def scores(self, group=None):
        """S[hi-lo, Nc] fp32 on the device for this shard's image block.

        With a process group (world > 1) and captions on the HOST, every rank uploads and packs
        only its 1/world share of the captions and the packed bf16 rows are all-gathered over
        NVLink (2.6 GB in total at COCO-5k) instead of every rank pulling all 5 GB over PCIe."""
        import torch.distributed as dist
        split = self.precision == "fp32"
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
            return scoring.mrsw_scores_packed(self.captions.packed, regions, tiles_dev, len(table), n_loc, self.Nc, out=S)
        shard_caps = (group is not None and self.world > 1 and not self.captions.is_cuda and dist.is_initialized())
        if (n_loc == 0 and not shard_caps) or self.Nc == 0:
            return S
        nr, nw = self.nr[lo:hi], self.nw
        d = self.captions.shape[2]
        Kp = ((d * (3 if split else 1) + _cabi.TILE_K - 1) // _cabi.TILE_K) * _cabi.TILE_K
        # ---- regions of this image block
        regions = tiles_dev = None
        n_tiles = 0
        if n_loc:
            Lr = 1 + int(nr.max())
            im_dev = _upload_rows(self.images, self.img_start + lo * self.img_step, self.img_step, n_loc, Lr)
            regions = scoring.pack_tokens(im_dev, nr, slot0=1, mode=2 if split else 0)
            _, table, _ = build_region_tiles(nr, self.clamp[lo:hi])
            n_tiles = len(table)
            tiles_dev = scoring._to_dev(table.view(np.int32).reshape(-1), dev) if n_tiles else None
        # ---- words: all captions (single rank / device-resident) or this rank's share + all-gather
        if shard_caps:
            # captions are processed in phases; inside a phase every rank uploads + packs its 1/world
            # share, the packed rows are all-gathered (NVLink) and the phase is scored while the next
            # phase is being prepared on a side stream
            P = self.caption_phases if self.Nc >= self.caption_phases * self.world * 64 else 1
            pb = [(p * self.Nc // P, (p + 1) * self.Nc // P) for p in range(P)]
            plans = []
            for c0, c1 in pb:
                spans = [tuple(c0 + x for x in shard_bounds(c1 - c0, self.world, r)) for r in range(self.world)]
                rows_max = max(int(nw[a:b].sum()) for a, b in spans)
                pad = ((rows_max + 2 * _cabi.TILE_M - 1) // (2 * _cabi.TILE_M)) * (2 * _cabi.TILE_M)
                plans.append((spans[self.rank], max(pad, 2 * _cabi.TILE_M)))
            pad_max = max(pad for _, pad in plans)
            words_buf = [torch.empty((self.world * pad_max, Kp), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            caps_buf = [torch.empty((self.world * pad_max,), dtype=torch.int32, device=dev) for _ in range(2)]
            mine_w = [torch.empty((pad_max, Kp), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            mine_c = [torch.empty((pad_max,), dtype=torch.int32, device=dev) for _ in range(2)]
            main = torch.cuda.current_stream()
            prep = torch.cuda.Stream()
            prep.wait_stream(main)
            freed = [None, None]
            for p, ((c0, c1), ((m0, m1), pad)) in enumerate(zip(pb, plans)):
                b = p & 1
                with torch.cuda.stream(prep):
                    if freed[b] is not None:
                        prep.wait_event(freed[b])
                    mw, mc = mine_w[b][:pad], mine_c[b][:pad]
                    mc.fill_(-1)
                    self._pack_caption_range(m0, m1, mw, mc, 0, split, dev, item_origin=c0)
                    wa, ca = words_buf[b][:self.world * pad], caps_buf[b][:self.world * pad]
                    dist.all_gather_into_tensor(wa, mw, group=group)      # gap rows are never scored (row_cap -1)
                    dist.all_gather_into_tensor(ca, mc, group=group)
                    ready = torch.cuda.Event()
                    ready.record(prep)
                main.wait_event(ready)
                if n_loc:
                    words = scoring.Packed(wa, self.world * pad, Kp, None, None, ca, 1 if split else 0)
                    scoring.mrsw_scores_packed(words, regions, tiles_dev, n_tiles, n_loc, c1 - c0, out=S[:, c0:c1])
                freed[b] = torch.cuda.Event()
                freed[b].record(main)
            return S
        else:
            # one rank owns all captions: score chunk k while chunk k+1 is uploaded (host sources)
            on_cpu = not self.captions.is_cuda
            chunk = self.caption_chunk if on_cpu else self.Nc
            # host sources: the first chunks are small (chunk/8, /4, /2) so that scoring starts after a few ms of
            # PCIe traffic instead of a whole chunk's worth; the copies stay ahead of the scoring from then on
            bounds, c0, size = [], 0, min(chunk, max(chunk // 8, 256)) if on_cpu else chunk
            while c0 < self.Nc:
                bounds.append((c0, min(self.Nc, c0 + size)))
                c0 += size
                size = min(chunk, 2 * size)
            main = torch.cuda.current_stream()
            if on_cpu:
                Lw_max = 1 + int(nw.max())
                stage = [torch.empty((chunk, Lw_max, d), dtype=torch.float32, device=dev) for _ in range(2)]
                copy_stream = torch.cuda.Stream()
                copy_stream.wait_stream(main)
                ready = [torch.cuda.Event() for _ in bounds]
                freed = [None, None]
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
                words = scoring.pack_tokens(cap_dev, nw[c0:c1], slot0=1, mode=1 if split else 0, want_row_item=True)
                scoring.mrsw_scores_packed(words, regions, tiles_dev, n_tiles, n_loc, c1 - c0, out=S[:, c0:c1])
                if on_cpu:
                    freed[bsel] = torch.cuda.Event()
                    freed[bsel].record(main)
            return S


def rank_both_directions(S, npts, img_off=0, n_images_total=None, k=50, group=None, gather_i2t=True, ops=ranking):
    """Exact i2t / t2i ranks and top-k from a shard's score block S[n_loc, Nc].

    Returns numpy arrays shaped like the reference's (alad/evaluation.py:166-167,255-256):
    ranks_i2t[npts], top1[npts], ranks_t2i[5*npts], topk[5*npts, k] (float64).
    `ops` provides rank_rows / col_gt / col_count / col_topk / topk_merge (the CUDA kernels by
    default; the gloo CPU tests plug the oracle in to exercise the exchange protocol)."""
    import torch.distributed as dist
    n_loc, Nc = S.shape
    dist_on = group is not None and dist.is_initialized() and dist.get_world_size(group) > 1
    Ni_total = n_images_total if n_images_total is not None else n_loc
    npts = min(npts, Ni_total)
    # ---------------- i2t: rows (queries = images < npts), gallery = all captions
    q_loc = max(0, min(n_loc, npts - img_off))
    rank_i, top1_i = ops.rank_rows(S[:q_loc], 5, img_off)
    # ---------------- t2i: columns (queries = captions < 5*npts), gallery = all images
    ncq = min(Nc, 5 * npts)
    Sq = S[:, :ncq]
    gt = torch.zeros(ncq, dtype=torch.float32, device=S.device)
    ops.col_gt(Sq, gt, 5, img_off)
    if dist_on:
        dist.all_reduce(gt, group=group)                       # every entry is owned by exactly one shard
    count = ops.col_count(Sq, gt, 5, img_off)
    cs, ci = ops.col_topk(Sq, k, img_off)
    ts, ti = ops.topk_merge(cs, ci)
    if dist_on:
        world = dist.get_world_size(group)
        # per-shard k-best lists -> every rank, then one merge of `world` sorted lists per caption
        gs = torch.empty((world * ts.shape[0], ts.shape[1]), dtype=ts.dtype, device=S.device)
        gi = torch.empty((world * ti.shape[0], ti.shape[1]), dtype=ti.dtype, device=S.device)
        dist.all_gather_into_tensor(gs, ts.contiguous(), group=group)       # output = concatenation along dim 0
        dist.all_gather_into_tensor(gi, ti.contiguous(), group=group)
        ts, ti = ops.topk_merge(gs.view(world, *ts.shape), gi.view(world, *ti.shape))
        # counts (summed over the shards) and the i2t results of every image block in one small gather
        per = (Ni_total + world - 1) // world
        small = torch.full((ncq + 2 * per,), -1, dtype=torch.int32, device=S.device)
        small[:ncq] = count
        small[ncq:ncq + q_loc] = rank_i
        small[ncq + per:ncq + per + q_loc] = top1_i
        gsm = torch.empty((world * (ncq + 2 * per),), dtype=torch.int32, device=S.device)
        dist.all_gather_into_tensor(gsm, small, group=group)
        gsm = gsm.view(world, ncq + 2 * per)
        count = gsm[:, :ncq].sum(dim=0, dtype=torch.int32)
        if gather_i2t:
            rank_i = gsm[:, ncq:ncq + per].reshape(-1)[:npts]
            top1_i = gsm[:, ncq + per:].reshape(-1)[:npts]
    return _to_host_f64(rank_i, top1_i, count, ti)


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


def recall_tuple(ranks):
    """(r1, r5, r10, medr, meanr) exactly as alad/evaluation.py:231-235."""
    n = ranks.size
    r1 = 100.0 * np.count_nonzero(ranks < 1) / n
    r5 = 100.0 * np.count_nonzero(ranks < 5) / n
    r10 = 100.0 * np.count_nonzero(ranks < 10) / n
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return r1, r5, r10, medr, meanr
