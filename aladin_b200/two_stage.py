"""Two-stage retrieval (BASELINE.json config 5): matching-head global cosine top-K shortlist, then
alignment-head re-rank of the shortlist.  Not a function of the reference; it composes the
reference's matching scores (alad/recall_auxiliary.py:30: ``ims @ caps.T`` on slot-0 vectors) with its alignment
scores (alad/loss.py:97-125, 'MrSw') and the oracle is that composition (tests/test_gpu_two_stage.py).

Stage 2 scores ONLY the shortlisted pairs (~2 % of the Ni x Nc block at K = 100), on the same tcgen05 mainloop as
the dense pass: ``alad_pairtile_build`` turns the shortlists into "pair tiles" -- 128 packed word rows of one or two
captions x up to 8 gathered image slots, each slot loaded by its own TMA box -- and ``alad_mrsw_scores_pairs`` runs
them (csrc/mrsw_fwd.cu, LIST mode).  Both directions share one pass: caption c is scored against its own t2i
shortlist and against the images whose i2t shortlist contains c.  Final order per query: shortlisted items by
alignment score, then the rest by stage-1 score (a ground truth outside the shortlist keeps its stage-1 rank).

Multi-GPU: image blocks as in retrieval.py (captions replicated).  Stage 1 exchanges the per-shard top-K candidates;
stage 2 scores the shortlisted pairs whose IMAGE is local, and one all-reduce of the [Nc, K] t2i scores (every
entry owned by one shard) completes the lists.  i2t lists are local to the image's owner."""
import ctypes as C

import numpy as np
import torch

from . import _cabi, ranking, retrieval, scoring

# Region rows per image slot = max scored regions of the shard, rounded up to SLOT_ALIGN.  TMA's 128-byte swizzle is a
# function of the shared-memory ADDRESS, so a slot box that starts between two 1024-byte swizzle atoms lands exactly
# where the MMA descriptor of the whole operand tile expects its rows (checked on hardware for both settings,
# tests/test_gpu_two_stage.py): no rounding, 7 slots of 34 rows per tile at COCO shape instead of 6 of 40.
SLOT_ALIGN = 1
# Packed region bytes per image block of the pair-list pass (see alad_pairtile_args.block_images): a gathered slot is
# re-used by ~Nc*K/Ni caption groups, so the block being swept has to stay in the 126 MB L2 next to the streamed words.
L2_BLOCK_BYTES = 48 << 20
# Load only the word rows a caption group has (104 of 128 for two 50-word captions) instead of the full 128-row box.
WORD_BOX = True
_GROUPS = {}
# diagnostics (tools/two_stage_probe.py): when a list, two_stage_retrieval appends (stage name, CUDA event) marks
stage_timeline = None


def _caption_groups(nw):
    """Host table of the M tiles (alad_caption_groups), memoised by content."""
    nw = np.ascontiguousarray(nw, dtype=np.int32)
    key = nw.tobytes()
    hit = _GROUPS.get(key)
    if hit is None:
        Nc = len(nw)
        row0 = np.empty(max(Nc, 1), np.int32)
        cap_lo = np.empty(Nc + 1, np.int32)
        cap_group = np.empty(max(Nc, 1), np.int32)
        n_g = _cabi.lib().alad_caption_groups(nw.ctypes.data, Nc, row0.ctypes.data, cap_lo.ctypes.data, cap_group.ctypes.data)
        _cabi.check(min(n_g, 0), "alad_caption_groups")
        if len(_GROUPS) >= 4:
            _GROUPS.clear()
        # longest group in packed word rows, rounded up to the 8-row swizzle atom: what a tile has to load
        rows = np.concatenate([[0], np.cumsum(nw, dtype=np.int64)])
        longest = int((rows[cap_lo[1:n_g + 1]] - rows[cap_lo[:n_g]]).max()) if n_g else 0
        box = min(_cabi.TILE_M, max(8, (longest + 7) // 8 * 8))
        hit = _GROUPS[key] = (n_g, row0[:max(n_g, 1)].copy(), cap_lo[:n_g + 1].copy(), cap_group, box)
    return hit


def slot_rows_for(nr):
    m = int(nr.max()) if len(nr) else 1
    s = max(_cabi.TILE_N // _cabi.PTILE_SLOTS, ((m + SLOT_ALIGN - 1) // SLOT_ALIGN) * SLOT_ALIGN)
    if s > _cabi.TILE_N:
        raise ValueError(f"an image has {m} scored regions; the kernel supports at most {_cabi.TILE_N}")
    return s


def pair_scores(words, regions, region_row_off, nr, clamp, nw, lists_t2i, lists_i2t, img_off, out=None):
    """S[n_loc, Nc] holding the MrSw scores of the listed pairs (other entries are undefined).
    lists_t2i [Nc, K] int32 global image ids (or None), lists_i2t [n_loc, Kc] int32 caption ids (or None)."""
    lib = _cabi.lib()
    dev = words.data.device
    n_loc, Nc = len(nr), len(nw)
    if out is None:
        out = torch.empty((n_loc, Nc), dtype=torch.float32, device=dev)
    if n_loc == 0 or Nc == 0 or regions is None:
        return out, None
    n_g, row0, cap_lo, cap_group, word_box = _caption_groups(nw)
    slot_rows = slot_rows_for(nr)
    slots = _cabi.TILE_N // slot_rows
    k1 = lists_t2i.shape[1] if lists_t2i is not None else 0
    k2 = lists_i2t.shape[1] if lists_i2t is not None else 0
    block = 0
    row_bytes = words.Kp * 2
    if regions.n_rows * row_bytes > L2_BLOCK_BYTES:
        per_image = max(regions.n_rows / n_loc * row_bytes, 1.0)
        block = max(32, int(L2_BLOCK_BYTES / per_image) // 32 * 32)
        if block >= n_loc:
            block = 0
    n_blocks = 1 if block == 0 else (n_loc + block - 1) // block
    capacity = (Nc * k1 + n_loc * k2) // slots + n_g * n_blocks + 1
    row0_d, cap_lo_d, cap_group_d, rr_d, nr_d, clamp_d = scoring._to_dev_group(
        [row0, cap_lo, cap_group, np.asarray(region_row_off, np.int32), np.asarray(nr, np.int32),
         np.asarray(clamp, np.uint8)], dev)
    ptiles = torch.empty((capacity, _cabi.PTILE_WORDS), dtype=torch.int32, device=dev)
    n_ptiles = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(lib.alad_pairtile_workspace_bytes(n_g, n_loc, block))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    a = _cabi.PairtileArgs(
        n_groups=n_g, group_row0=row0_d.data_ptr(), group_cap_lo=cap_lo_d.data_ptr(), cap_group=cap_group_d.data_ptr(), Nc=Nc,
        lists_t2i=lists_t2i.data_ptr() if k1 else None, k_t2i=k1, lists_i2t=lists_i2t.data_ptr() if k2 else None, k_i2t=k2,
        img_off=img_off, n_loc=n_loc, region_row=rr_d.data_ptr(), nr=nr_d.data_ptr(), clamp=clamp_d.data_ptr(),
        slot_rows=slot_rows, block_images=block, ptiles=ptiles.data_ptr(), capacity=capacity, n_ptiles=n_ptiles.data_ptr(),
        workspace=ws.data_ptr(), workspace_bytes=nbytes)
    _cabi.check(lib.alad_pairtile_build(C.byref(a), _cabi.stream_ptr()), "alad_pairtile_build")
    b = _cabi.MrswPairsArgs(
        words=words.data.data_ptr(), n_word_rows=words.n_rows, regions=regions.data.data_ptr(), n_region_rows=regions.n_rows,
        Kp=words.Kp, row_cap=words.row_item.data_ptr(), ptiles=ptiles.data_ptr(), n_ptiles=n_ptiles.data_ptr(),
        max_ptiles=capacity, slot_rows=slot_rows, S=out.data_ptr(), ldS=max(out.stride(0), Nc), Ni=n_loc, Nc=Nc,
        transpose_out=0, num_ctas=0, word_box_rows=word_box if WORD_BOX else 0)
    if scoring.kernel_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cabi.check(lib.alad_mrsw_scores_pairs(C.byref(b), _cabi.stream_ptr()), "alad_mrsw_scores_pairs")
    if scoring.kernel_timeline is not None:
        e1.record()
        scoring.kernel_timeline.append((e0, e1, n_ptiles, slots, words.Kp))
    return out, n_ptiles


def gather_list_scores(S, ids, by_column, img_off, nr_d, nw_d):
    Q, k = ids.shape
    out = torch.empty((Q, k), dtype=torch.float32, device=S.device)
    Ni, Nc = S.shape
    _cabi.check(_cabi.lib().alad_gather_list_scores(S.data_ptr(), max(S.stride(0), Nc), Ni, Nc, ids.data_ptr(), Q, k,
                                                    1 if by_column else 0, img_off, nr_d.data_ptr(), nw_d.data_ptr(),
                                                    out.data_ptr(), _cabi.stream_ptr()), "alad_gather_list_scores")
    return out


def list_rerank(scores, ids, q_off, gt_mul, gt_div, gt_n, fallback, want_order=True):
    """(rank[Q] int32, order[Q, k] int32): ids re-ordered by (score desc, list position desc) and the position of the best
    ground-truth id inside that order (fallback[q] when no ground truth is listed)."""
    Q, k = ids.shape
    rank = torch.empty(Q, dtype=torch.int32, device=ids.device)
    order = torch.empty((Q, k), dtype=torch.int32, device=ids.device) if want_order else None
    _cabi.check(_cabi.lib().alad_list_rerank(scores.data_ptr(), ids.data_ptr(), Q, k, q_off, gt_mul, gt_div, gt_n,
                                             fallback.data_ptr() if fallback is not None else None, rank.data_ptr(),
                                             order.data_ptr() if order is not None else None, _cabi.stream_ptr()),
                "alad_list_rerank")
    return rank, order


def two_stage_retrieval(images, captions, img_lens, cap_lens, shortlist=100, precision=None, return_ranks=False,
                        group="auto", return_details=False):
    """images/captions: evaluation containers [5*Ni, S, d] / [Nc, S, d] (slot 0 = global vector,
    alad/evaluation.py:119-130).  Returns ((r1,r5,r10,medr,meanr) i2t, (...) t2i) and optionally the rank arrays.
    Under torch.distributed (world > 1; ``group='auto'`` picks the default group) the gallery images are sharded."""
    import torch.distributed as dist
    if group == "auto":
        group = dist.group.WORLD if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    def mark(name):
        if stage_timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stage_timeline.append((name, ev))

    mark("start")
    Ni = images.shape[0] // 5
    Nc = captions.shape[0]
    k = min(shortlist, Ni)
    kc = min(shortlist, Nc)
    precision = precision or scoring.get_precision()
    gal = retrieval.AlignmentGallery(images, captions, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5,
                                     precision=precision, world=world, rank=rank,
                                     bounds=[retrieval.shard_bounds(Ni, world, r) for r in range(world)])
    lo, hi = gal.lo, gal.hi
    n_loc = hi - lo
    # ---- stage 1: global-vector scores (alad/recall_auxiliary.py:30), shortlists and stage-1 ranks
    ims_g = scoring._require_cuda(images[0::5][lo:hi, 0, :], "images")
    caps_g = scoring._require_cuda(captions[:, 0, :], "captions")
    dev = caps_g.device
    mark("gallery")
    M = scoring.dot_scores(ims_g, caps_g, precision="fp32")                 # [n_loc, Nc]
    mark("stage1_gemm")
    rank1_i2t, _, rank1_t2i, _, short_t2i, _ = retrieval.rank_device(M, Ni, img_off=lo, n_images_total=Ni, k=k, group=group,
                                                                     bounds=gal.bounds)
    short_t2i = short_t2i.contiguous()                                      # [Nc, k] global image ids, best first
    mark("stage1_rank_t2i_lists")
    Mt = scoring.dot_scores(caps_g, ims_g, precision="fp32")                # [Nc, n_loc] (columns = local images)
    cs, ci = ranking.col_topk(Mt, kc)
    _, short_i2t = ranking.topk_merge(cs, ci)
    short_i2t = short_i2t.contiguous()                                      # [n_loc, kc] caption ids per local image
    mark("stage1_i2t_lists")
    # ---- stage 2: alignment scores of the shortlisted pairs only
    words, regions, region_row_off = gal.packed_operands()
    mark("pack_tokens")
    nr, nw, clamp = gal.nr[lo:hi], gal.nw, gal.clamp[lo:hi]
    S, n_ptiles = pair_scores(words, regions, region_row_off, nr, clamp, nw, short_t2i, short_i2t, lo)
    mark("pair_tiles_and_scores")
    nr_d, nw_d = scoring._to_dev_group([np.asarray(nr, np.int32), np.asarray(nw, np.int32)], dev)
    sc_t2i = gather_list_scores(S, short_t2i, True, lo, nr_d, nw_d)         # 0 where the image is another shard's
    if world > 1:
        dist.all_reduce(sc_t2i, group=group)
    r2_t2i, order_t2i = list_rerank(sc_t2i, short_t2i, 0, 1, 5, 1, rank1_t2i, want_order=return_details)
    sc_i2t = gather_list_scores(S, short_i2t, False, lo, nr_d, nw_d)
    r2_i2t_loc, order_i2t = list_rerank(sc_i2t, short_i2t, lo, 5, 1, 5, rank1_i2t[lo:hi].contiguous(), want_order=return_details)
    if world > 1:
        per = max(b - a for a, b in gal.bounds)
        mine = torch.full((per,), -1, dtype=torch.int32, device=dev)
        mine[:n_loc] = r2_i2t_loc
        allr = torch.empty((world * per,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allr, mine, group=group)
        r2_i2t = torch.cat([allr[r * per:r * per + (b - a)] for r, (a, b) in enumerate(gal.bounds)])
    else:
        r2_i2t = r2_i2t_loc
    mark("rerank")
    ri, rt = retrieval._to_host_f64(r2_i2t, r2_t2i)
    mark("to_host")
    out = (retrieval.recall_tuple(ri), retrieval.recall_tuple(rt))
    if return_details:
        return out, dict(ranks_i2t=ri, ranks_t2i=rt, order_t2i=order_t2i, order_i2t=order_i2t, short_t2i=short_t2i,
                         short_i2t=short_i2t, scores_t2i=sc_t2i, scores_i2t=sc_i2t, n_ptiles=n_ptiles, S=S)
    return (out, (ri, rt)) if return_ranks else out
