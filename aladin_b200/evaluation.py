"""Drop-ins for ``alad.evaluation.i2t`` / ``t2i`` (alad/evaluation.py:158-327).

Same signatures, same return tuples (7 metrics, optionally (ranks, top1) / (ranks, top50) as
float64 numpy arrays).  When ``sim_function`` is (or closes over) an
``aladin_b200.loss.AlignmentContrastiveLoss`` with aggregation 'MrSw' -- which is exactly
what alad/test.py:258-264 and alad/train.py:493-500 build -- the whole Ni x Nc score block is
computed once on the GPU and both directions are ranked from it (the other pooling modes of the
drop-in criterion get the same treatment through one criterion call on the whole gallery); ``sim_function=None`` is the
global-vector path on slot 0 (evaluation.py:195-197,284-286); any other callable is invoked
per query like the reference does, and only the ranking runs in our kernels."""
import weakref

import numpy as np
import torch

from . import retrieval, scoring

_cache = {}
# Score matrices larger than this (or than 40 % of the free device memory) are never materialised: the fused path then ranks
# block by block (retrieval.streaming_ranks); COCO-5k needs 0.5 GB.
STREAM_SCORE_BYTES = 32 << 30
# verdict of the closure probe per sim_function object (i2t and t2i of one evaluation hand over the same closure)
_probe_memo = weakref.WeakKeyDictionary()


def fused_sim_function(scorer):
    """The ``sim_function`` closure of alad/test.py:259-263 over a drop-in criterion, TAGGED so that i2t / t2i take
    the fused one-pass path without probing it: ``fn(img, cap, img_len, cap_len)`` = scores of the criterion."""
    def alignment_sim_fn(img, cap, img_len, cap_len):
        with torch.no_grad():
            return scorer(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)
    alignment_sim_fn.alad_scorer = scorer
    return alignment_sim_fn


def _probe_ok(fn, scorer, images, captions, img_lens, cap_lens):
    """Does ``fn`` return exactly the criterion's scores?  One query image x up to 8 captions through both; a closure
    that post-processes the scores (adds the matching term, rescales, combines two criteria) differs and is then
    invoked per query like the reference does."""
    from .gallery import DeviceContainer
    if isinstance(images, DeviceContainer):
        return True                      # packed containers have no raw tokens a foreign callable could consume
    n = min(8, captions.shape[0])
    if images.shape[0] == 0 or n == 0:
        return True
    try:
        with torch.no_grad():
            im = images[0:1].cuda()
            cap = captions[:n].cuda()
            a = fn(im, cap, [img_lens[0]], list(cap_lens[:n]))
            b = scorer(im, cap, [img_lens[0]], list(cap_lens[:n]), return_loss=False, return_similarity_mat=True)
        a, b = torch.as_tensor(a).float().reshape(-1).cpu(), b.float().reshape(-1).cpu()
        return a.shape == b.shape and bool(torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)))
    except Exception:
        return False


def _find_scorer(fn, probe=None):
    """The drop-in criterion behind ``sim_function`` when the fused path may replace the callable:
    the criterion instance itself, a callable tagged by ``fused_sim_function`` (attribute ``alad_scorer``), or a bound
    method / closure over a criterion that the probe (``probe`` = (images, captions, img_lens, cap_lens)) shows to
    return the criterion's scores unchanged.  Anything else -> None (per-query callback path, like the reference)."""
    from .loss import AlignmentContrastiveLoss

    def valid(c):
        return isinstance(c, AlignmentContrastiveLoss) and c.aggregation in c.SUPPORTED

    if fn is None:
        return None
    if valid(fn):
        return fn
    tagged = getattr(fn, "alad_scorer", None)
    if valid(tagged):
        return tagged
    cands = [getattr(fn, "__self__", None)]
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            cands.append(cell.cell_contents)
        except ValueError:
            pass
    for c in cands:
        if valid(c):
            if probe is None:
                return c
            memo_key = (id(c), c.aggregation, getattr(c, "precision", None))
            try:
                hit = _probe_memo.get(fn)
            except TypeError:                # not weak-referenceable: probe every time
                hit = None
            if hit is None or hit[0] != memo_key:
                hit = (memo_key, _probe_ok(fn, c, *probe))
                try:
                    _probe_memo[fn] = hit
                except TypeError:
                    pass
            return c if hit[1] else None
    return None


def _key(images, captions, lkey, mode, precision):
    from .gallery import DeviceContainer
    if isinstance(images, DeviceContainer):
        return (id(images), id(captions), images.packed.data.data_ptr(), captions.packed.data.data_ptr(), mode, precision)
    return (images.data_ptr(), captions.data_ptr(), tuple(images.shape), tuple(captions.shape), images._version,
            captions._version, lkey, mode, precision)


def _callback_scores(images, captions, img_lens, cap_lens, sim_function, cap_batches):
    """Reference behaviour for arbitrary callables: one sim_function call per query image and
    gallery chunk (evaluation.py:199-210); results land in a device matrix for our ranking."""
    N = images.shape[0]
    Ni = N // 5
    Nc = captions.shape[0]
    # the gallery chunks cover EVERY caption (the reference's `shape[0] // batches` silently drops the remainder of a
    # non-divisible split: evaluation.py:173,262); one matrix serves i2t and t2i
    per = max(1, -(-Nc // max(int(cap_batches), 1)))
    spans = [(c0, min(Nc, c0 + per)) for c0 in range(0, Nc, per)]
    S = torch.empty((Ni, Nc), dtype=torch.float32, device="cuda")
    caps_dev = [captions[c0:c1].cuda() for c0, c1 in spans]
    for i in range(Ni):
        im = images[5 * i].reshape(1, images.shape[1], images.shape[2]).cuda()
        for (c0, c1), cd in zip(spans, caps_dev):
            d = sim_function(im, cd, [img_lens[5 * i]], cap_lens[c0:c1])
            S[i, c0:c1] = d.reshape(-1).float().cuda()
    return S


def _block_scores(images, captions, img_lens, cap_lens, scorer):
    """Any pooling mode of the drop-in criterion other than 'MrSw' ('MrAVGw', 'MwSr', 'symm', 'sum', 'mean',
    'scan-sentences'): ONE criterion call on the distinct gallery images x all captions instead of one call per
    query and gallery chunk (evaluation.py:199-210, 288-300) -- every pair's score is independent of how the
    calls are batched, so the block equals the matrix the per-query loop assembles."""
    with torch.no_grad():
        ims = images[0::5].cuda()
        caps = captions.cuda()
        return scorer(ims, caps, list(img_lens[0::5]), list(cap_lens), return_loss=False, return_similarity_mat=True).float()


def clear_cache():
    """Forget the score block kept from the previous i2t/t2i call (and the closure-probe verdicts)."""
    _cache.clear()
    _probe_memo.clear()


def _evict(key):
    """weakref.finalize callback: the inputs of the cached block died -> release its HBM (0.5 GB at COCO-5k)."""
    if _cache.get("key") == key:
        _cache.clear()


def _dist_state():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_world_size(), dist.get_rank(), dist.group.WORLD
    return 1, 0, None


def _retrieve(images, captions, img_lens, cap_lens, sim_function, batches):
    """Scores + both directions' ranks for the whole gallery, computed once per distinct input.

    Returns dict(S, img_off, ranks_i2t, top1, ranks_t2i, top50).  Under torch.distributed
    (world > 1) the fused path shards the gallery images across ranks (retrieval.py)."""
    scorer = _find_scorer(sim_function, probe=(images, captions, img_lens, cap_lens))
    mode = "global" if sim_function is None else ("callback" if scorer is None else
                                                   "fused" if scorer.aggregation == "MrSw" else "block")
    from .gallery import DeviceContainer
    if mode == "block" and (_dist_state()[0] > 1 or isinstance(images, DeviceContainer)):
        mode = "callback"                     # sharding and packed containers exist for the 'MrSw' kernel only
    precision = getattr(scorer, "precision", None) or scoring.get_precision()
    mode_key = mode if scorer is None else f"{mode}:{scorer.aggregation}"
    lkey = retrieval.lens_key(img_lens, cap_lens) if mode != "callback" else None     # one pass over the python lists per call
    key = _key(images, captions, lkey, mode_key, precision) if mode != "callback" else None
    if (key is not None and _cache.get("key") == key and _cache["refs"][0]() is images
            and _cache["refs"][1]() is captions):
        return _cache["res"]
    Ni = images.shape[0] // 5
    world, rank, group = (1, 0, None)
    img_off, bounds = 0, None
    if isinstance(images, DeviceContainer) and mode == "callback":
        raise TypeError("DeviceContainer galleries hold packed tokens only: pass sim_function=None or a closure over "
                        "aladin_b200.loss.AlignmentContrastiveLoss('MrSw')")
    if mode == "global":
        if isinstance(images, DeviceContainer):
            S = scoring.dot_scores(images[:, 0, :][0::5], captions[:, 0, :], precision=precision)
        else:
            S = scoring.dot_scores(images[0::5][:, 0, :], captions[:, 0, :], precision=precision)
    elif mode == "fused" and _dist_state()[0] == 1 and not isinstance(images, DeviceContainer) and _too_large(Ni, captions.shape[0]):
        # the [Ni, Nc] block does not fit: ranks and lists block by block, S is never materialised
        ri, t1, rt, tk = retrieval.streaming_ranks(images, captions, img_lens, cap_lens, Ni, img_start=0, img_step=5,
                                                   precision=precision, block_images=_stream_block(captions.shape[0]),
                                                   k=min(50, Ni))
        res = dict(S=None, img_off=0, world=1, ranks_i2t=ri, top1=None, ranks_t2i=rt, top50=None, lists=(t1, tk))
        _store(key, res, images, captions)
        return res
    elif mode == "fused":
        world, rank, group = _dist_state()
        gal = retrieval.AlignmentGallery(images, captions, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5,
                                         precision=precision, world=world, rank=rank, lkey=lkey)
        S = gal.scores(group=group)
        img_off, bounds = gal.lo, gal.bounds
    elif mode == "block":
        S = _block_scores(images, captions, img_lens, cap_lens, scorer)
    else:
        S = _callback_scores(images, captions, img_lens, cap_lens, sim_function, batches)
    k = min(50, Ni)
    # ranks to the host now; the top-1 / top-50 lists stay on the device until a caller asks for them (return_ranks)
    ri, t1, rt, tk = retrieval.rank_both_directions(S, Ni, img_off=img_off, n_images_total=Ni, k=k, group=group, bounds=bounds,
                                                    lists_to_host=False)
    res = dict(S=S, img_off=img_off, world=world, ranks_i2t=ri, top1=t1, ranks_t2i=rt, top50=tk)
    _store(key, res, images, captions)
    return res


def _store(key, res, images, captions):
    if key is None:
        return
    # the key holds addresses: a hit is valid only while the very same input objects are alive (a freed
    # tensor's address, shape and version counter can all recur, e.g. the next epoch's validation embeddings)
    refs = (weakref.ref(images), weakref.ref(captions))
    _cache.clear()
    _cache.update(key=key, res=res, refs=refs)
    for obj in (images, captions):       # the block dies with its inputs, not with the next call
        weakref.finalize(obj, _evict, key)


def _too_large(Ni, Nc):
    need = 4 * Ni * Nc
    if need > STREAM_SCORE_BYTES:
        return True
    try:
        free, _ = torch.cuda.mem_get_info()
    except Exception:
        return False
    return need > 0.4 * free


def _stream_block(Nc):
    """Images per block of the streaming path: a ~1 GiB score buffer, at least 64 images."""
    return max(64, int((1 << 30) // max(4 * Nc, 1)))


def _lists(res):
    """(top1, top50) host arrays of a cached block, fetched from the device on first use."""
    if res.get("lists") is None:
        res["lists"] = retrieval.lists_host(res["top1"], res["top50"])
    return res["lists"]


def _ndcg(ndcg_scorer, res, npts, fold_index, retrieval_kind):
    """NDCG hooks (evaluation.py:225-228,310-313): need the full order per query."""
    if res["world"] > 1 or res["S"] is None:
        raise NotImplementedError("NDCG scoring needs the whole score block on one device")
    S = res["S"]
    n = npts if retrieval_kind == "sentence" else 5 * npts
    rougel, spice = np.zeros(n), np.zeros(n)
    for q in range(n):
        d = S[q] if retrieval_kind == "sentence" else S[:, q]
        # same total order as the ranking kernels and as numpy.argsort(d)[::-1]: score descending, HIGHER index first
        # on exact ties (stable descending sort of the reversed vector, mapped back)
        inds = (d.numel() - 1 - torch.argsort(d.flip(0), descending=True, stable=True)).cpu().numpy()
        rougel[q], spice[q] = ndcg_scorer.compute_ndcg(npts, q, inds.astype(int), fold_index=fold_index,
                                                       retrieval=retrieval_kind).values()
    return rougel, spice


def _check_measure(measure):
    if measure == "order":
        raise NotImplementedError("measure='order' (order embeddings) is outside the ported path: no shipped "
                                  "ALADIN config selects it (configs/*.yaml: measure: 'dot')")


def i2t(images, captions, img_lenghts, cap_lenghts, npts=None, return_ranks=False, ndcg_scorer=None, fold_index=0,
        measure='dot', sim_function=None, cap_batches=1):
    """Images->Text (image annotation).  Mirrors alad/evaluation.py:158-241."""
    _check_measure(measure)
    if npts is None:
        npts = images.shape[0] // 5
    res = _retrieve(images, captions, img_lenghts, cap_lenghts, sim_function, cap_batches)
    ranks = res["ranks_i2t"][:npts]
    if ndcg_scorer is not None:
        _ndcg(ndcg_scorer, res, npts, fold_index, "sentence")
    metrics = retrieval.recall_tuple(ranks) + (0, 0)
    return (metrics, (ranks.copy(), _lists(res)[0][:npts].copy())) if return_ranks else metrics


def t2i(images, captions, img_lenghts, cap_lenghts, npts=None, return_ranks=False, ndcg_scorer=None, fold_index=0,
        measure='dot', sim_function=None, im_batches=1):
    """Text->Images (image search).  Mirrors alad/evaluation.py:244-327."""
    _check_measure(measure)
    if npts is None:
        npts = images.shape[0] // 5
    if images.shape[0] // 5 < 50:
        raise ValueError("t2i keeps the 50 best images per caption (evaluation.py:257,308): need >= 50 gallery images")
    res = _retrieve(images, captions, img_lenghts, cap_lenghts, sim_function, im_batches)
    ranks = res["ranks_t2i"][:5 * npts]
    if ndcg_scorer is not None:
        _ndcg(ndcg_scorer, res, npts, fold_index, "image")
    metrics = retrieval.recall_tuple(ranks) + (0, 0)
    # the caller owns what it gets (the block stays cached): copies only when the arrays are asked for (top50: 10 MB at COCO-5k)
    return (metrics, (ranks.copy(), _lists(res)[1][:5 * npts].copy())) if return_ranks else metrics
