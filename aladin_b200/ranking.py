"""Device-side ranking ops over an image-major score matrix S[Ni,Nc] (thin wrappers over
the C ABI).  Replace numpy.argsort/where of alad/evaluation.py:213-223,303-308 and
alad/recall_auxiliary.py:34-56."""
import torch

from . import _cabi


def _check_S(S):
    assert S.is_cuda and S.dtype == torch.float32 and S.dim() == 2 and S.stride(1) == 1
    assert S.shape[0] <= 1 or S.stride(0) >= S.shape[1], "rows of S overlap (expanded / broadcast matrix): call .contiguous()"
    return S.shape[0], S.shape[1], max(S.stride(0), S.shape[1])


def rank_rows(S, group=5, img_off=0):
    """i2t: (rank[Ni] int32, top1[Ni] int32) -- evaluation.py:213-223."""
    Ni, Nc, ld = _check_S(S)
    rank = torch.empty(Ni, dtype=torch.int32, device=S.device)
    top1 = torch.empty(Ni, dtype=torch.int32, device=S.device)
    _cabi.check(_cabi.lib().alad_rank_rows(S.data_ptr(), ld, Ni, Nc, group, img_off, rank.data_ptr(), top1.data_ptr(),
                                            _cabi.stream_ptr()), "alad_rank_rows")
    return rank, top1


def col_gt(S, gt, group=5, img_off=0):
    """t2i: write the ground-truth score of every caption whose image this shard owns into gt[Nc]."""
    Ni, Nc, ld = _check_S(S)
    assert gt.is_cuda and gt.dtype == torch.float32 and gt.numel() == Nc and gt.is_contiguous()
    _cabi.check(_cabi.lib().alad_col_gt(S.data_ptr(), ld, Ni, Nc, group, img_off, gt.data_ptr(), _cabi.stream_ptr()),
                "alad_col_gt")
    return gt


def col_count(S, gt, group=5, img_off=0):
    """t2i: count[Nc] int32 = local images ordered ahead of each caption's ground truth."""
    Ni, Nc, ld = _check_S(S)
    count = torch.empty(Nc, dtype=torch.int32, device=S.device)
    _cabi.check(_cabi.lib().alad_col_count(S.data_ptr(), ld, Ni, Nc, group, img_off, gt.data_ptr(), count.data_ptr(),
                                            _cabi.stream_ptr()), "alad_col_count")
    return count


def col_topk(S, k, img_off=0, splits=None):
    """t2i: sorted per-caption candidates (score[P,Nc,k], global image idx[P,Nc,k]).

    Default (splits=None): the threshold-select kernels -- two sweeps of S, P = 1, already final for this
    shard.  An explicit `splits` runs the shared-memory heap kernel over that many row slices (P = splits
    lists to be merged with topk_merge); kept for A/B and as the overflow path of the default."""
    Ni, Nc, ld = _check_S(S)
    if splits is None:
        cs = torch.empty((1, Nc, k), dtype=torch.float32, device=S.device)
        ci = torch.empty((1, Nc, k), dtype=torch.int32, device=S.device)
        nbytes = _cabi.lib().alad_col_topk_select_workspace_bytes(Ni, Nc, k)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=S.device)
        _cabi.check(_cabi.lib().alad_col_topk_select(S.data_ptr(), ld, Ni, Nc, k, img_off, cs.data_ptr(), ci.data_ptr(),
                                                      ws.data_ptr(), _cabi.stream_ptr()), "alad_col_topk_select")
        return cs, ci
    splits = max(1, min(splits, (Ni + 63) // 64))
    cs = torch.empty((splits, Nc, k), dtype=torch.float32, device=S.device)
    ci = torch.empty((splits, Nc, k), dtype=torch.int32, device=S.device)
    _cabi.check(_cabi.lib().alad_col_topk(S.data_ptr(), ld, Ni, Nc, k, img_off, splits, cs.data_ptr(), ci.data_ptr(),
                                           _cabi.stream_ptr()), "alad_col_topk")
    return cs, ci


def rank_fused(S, k, img_off=0, q_rows=None, q_cols=None, gt=None, count=True, group=5):
    """Both directions of one score block in two sweeps of S (alad_rank_fused): i2t (rank, top1) of rows [0, q_rows)
    over all captions, t2i top-k (score[q_cols, k], global image idx[q_cols, k]) of captions [0, q_cols) over the block's
    rows and -- with count=True -- their "images ahead of the ground truth" counts against `gt` (default: the block's
    own entries, i.e. a block that holds every image).  Returns (rank, top1, count or None, topk_score, topk_idx);
    identical to rank_rows + col_gt + col_count + col_topk on the same block."""
    Ni, Nc, ld = _check_S(S)
    q_rows = Ni if q_rows is None else q_rows
    q_cols = Nc if q_cols is None else q_cols
    dev = S.device
    rank = torch.empty(q_rows, dtype=torch.int32, device=dev)
    top1 = torch.empty(q_rows, dtype=torch.int32, device=dev)
    cnt = torch.empty(q_cols, dtype=torch.int32, device=dev) if count else None
    ts = torch.empty((q_cols, k), dtype=torch.float32, device=dev)
    ti = torch.empty((q_cols, k), dtype=torch.int32, device=dev)
    if gt is not None:
        assert count and gt.is_cuda and gt.dtype == torch.float32 and gt.numel() == q_cols and gt.is_contiguous()
    nbytes = _cabi.lib().alad_rank_fused_workspace_bytes(Ni, q_rows, q_cols, k)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    _cabi.check(_cabi.lib().alad_rank_fused(S.data_ptr(), ld, Ni, Nc, group, img_off, q_rows, q_cols, k,
                                             gt.data_ptr() if gt is not None else None, rank.data_ptr(), top1.data_ptr(),
                                             cnt.data_ptr() if count else None, ts.data_ptr(), ti.data_ptr(), ws.data_ptr(),
                                             _cabi.stream_ptr()), "alad_rank_fused")
    if count and gt is None:
        _cabi.launch_count["kernels"] += 1          # the block's own ground-truth gather
    return rank, top1, cnt, ts, ti


def topk_merge(cand_score, cand_idx):
    """Merge P sorted candidate lists per caption: [P,Nc,k] -> ([Nc,k], [Nc,k])."""
    P, Nc, k = cand_score.shape
    if P == 1:                                   # a single sorted list per caption is already the result
        return cand_score[0], cand_idx[0]
    assert cand_score.is_contiguous() and cand_idx.is_contiguous() and cand_idx.shape == cand_score.shape
    os_ = torch.empty((Nc, k), dtype=torch.float32, device=cand_score.device)
    oi = torch.empty((Nc, k), dtype=torch.int32, device=cand_score.device)
    _cabi.check(_cabi.lib().alad_topk_merge(cand_score.data_ptr(), cand_idx.data_ptr(), P, Nc, k, os_.data_ptr(),
                                             oi.data_ptr(), _cabi.stream_ptr()), "alad_topk_merge")
    return os_, oi


def t2i_rank_topk(S, k, group=5):
    """Single-shard t2i: (rank[Nc] int32, topk[Nc,k] int32)."""
    _, _, count, _, idx = rank_fused(S, k, q_rows=0, group=group)
    return count, idx
