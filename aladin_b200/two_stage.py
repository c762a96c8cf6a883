"""Two-stage retrieval (BASELINE.json config 5): matching-head global cosine top-K shortlist,
then alignment-head re-rank of the shortlist.  Not part of the reference; the
oracle is the composition of reference functions (tests/test_gpu_two_stage.py).

B200-first choice: on this hardware the dense fused alignment pass over ALL pairs (0.35 s at
COCO-5k on one GPU) is cheaper than any gather of the 2 % shortlisted pairs onto CUDA cores, so
stage 2 scores everything with the tcgen05 kernel and the shortlist is applied as a mask
(non-shortlisted pairs -> -inf) before the exact ranking kernels.  Final order per query:
shortlisted items by alignment score, then the rest by stage-1 score."""
import numpy as np
import torch

from . import _cabi, ranking, retrieval, scoring


def _scatter(S, idx, by_column, img_off=0):
    Ni, Nc = S.shape
    S2 = torch.empty_like(S)
    n_lists, k = idx.shape
    _cabi.check(_cabi.lib().alad_shortlist_scatter(S.data_ptr(), max(S.stride(0), Nc), S2.data_ptr(), max(S2.stride(0), Nc),
                                                   Ni, Nc, idx.data_ptr(), n_lists, k, 1 if by_column else 0, img_off,
                                                   _cabi.stream_ptr()), "alad_shortlist_scatter")
    return S2


def two_stage_retrieval(images, captions, img_lens, cap_lens, shortlist=100, precision=None, return_ranks=False):
    """images/captions: evaluation containers [5*Ni, S, d] / [Nc, S, d] (slot 0 = global vector,
    alad/evaluation.py:119-130).  Returns ((r1,r5,r10,medr,meanr) i2t, (...) t2i) and optionally the
    rank arrays.  Single GPU."""
    Ni = images.shape[0] // 5
    Nc = captions.shape[0]
    k = min(shortlist, Ni)
    kc = min(shortlist, Nc)
    precision = precision or scoring.get_precision()
    # ---- stage 1: global-vector scores (alad/recall_auxiliary.py:30) and shortlists
    ims_g = scoring._require_cuda(images[0::5][:, 0, :], "images")
    caps_g = scoring._require_cuda(captions[:, 0, :], "captions")
    M = scoring.dot_scores(ims_g, caps_g, precision="fp32")                 # [Ni, Nc]
    Mt = scoring.dot_scores(caps_g, ims_g, precision="fp32")                # [Nc, Ni] (columns = images)
    cs, ci = ranking.col_topk(M, k)
    _, short_t2i = ranking.topk_merge(cs, ci)                               # [Nc, k] images per caption
    cs, ci = ranking.col_topk(Mt, kc)
    _, short_i2t = ranking.topk_merge(cs, ci)                               # [Ni, kc] captions per image
    rank1_i2t, _ = ranking.rank_rows(M, 5, 0)
    gt = torch.zeros(Nc, dtype=torch.float32, device=M.device)
    ranking.col_gt(M, gt, 5, 0)
    rank1_t2i = ranking.col_count(M, gt, 5, 0)
    # ---- stage 2: alignment scores, shortlist applied as a mask, exact ranks
    gal = retrieval.AlignmentGallery(images, captions, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5,
                                     precision=precision)
    S = gal.scores()
    S_t2i = _scatter(S, short_t2i.contiguous(), by_column=True)
    S_i2t = _scatter(S, short_i2t.contiguous(), by_column=False)
    r2_i2t, top1 = ranking.rank_rows(S_i2t, 5, 0)
    gt2 = torch.full((Nc,), float("-inf"), dtype=torch.float32, device=S.device)
    ranking.col_gt(S_t2i, gt2, 5, 0)
    r2_t2i = ranking.col_count(S_t2i, gt2, 5, 0)
    # ground truth inside the shortlist <=> its masked score is finite; otherwise keep the stage-1 rank
    best_gt = S_i2t.view(Ni, Nc)[torch.arange(Ni, device=S.device).repeat_interleave(5),
                                 torch.arange(5 * Ni, device=S.device)].view(Ni, 5).max(dim=1).values
    ranks_i2t = torch.where(torch.isfinite(best_gt), r2_i2t, rank1_i2t)
    ranks_t2i = torch.where(torch.isfinite(gt2), r2_t2i, rank1_t2i)
    ri = ranks_i2t.cpu().numpy().astype(np.float64)
    rt = ranks_t2i.cpu().numpy().astype(np.float64)
    out = (retrieval.recall_tuple(ri), retrieval.recall_tuple(rt))
    return (out, (ri, rt)) if return_ranks else out
