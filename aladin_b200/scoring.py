"""Device-side scoring ops (thin wrappers over the C ABI): token packing, the fused
MrSw alignment scores and the global-vector dot scores.

Reference call sites replaced (mesnico/ALADIN):
  alad/loss.py:79-125  AlignmentContrastiveLoss.forward (aggregation 'MrSw')
  alad/loss.py:8-18    dot_sim / cosine_sim
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .tiling import (build_region_tiles, exclusive_cumsum, gemm_tiles, padded_rows, round_up, valid_counts)

PRECISIONS = ("bf16", "fp32")
# bench.py sets this to a list to collect (start_event, end_event, pairs, Kp) of every scoring launch
kernel_timeline = None
_precision = "bf16"


def set_precision(mode):
    """'bf16': bf16 operands, fp32 accumulate (<= 1e-2 abs on scores).
    'fp32': split-precision hi/lo bf16 operands (3 tensor-core products per dot product,
    fp32-grade results, <= 1e-4 rel)."""
    global _precision
    if mode not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}")
    _precision = mode


def get_precision():
    return _precision


def _require_cuda(x, name):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not torch.cuda.is_available():
        raise _cabi.AladError("aladin_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    if x.device.type != "cuda":
        x = x.cuda(non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() >= 1 and x.stride(-1) != 1:
        x = x.contiguous()
    return x


def _to_dev(arr, device):
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True)


class Packed:
    """Dense K-major bf16 token rows on the device."""
    __slots__ = ("data", "n_rows", "Kp", "counts", "row_off", "row_item", "mode")

    def __init__(self, data, n_rows, Kp, counts, row_off, row_item, mode):
        self.data, self.n_rows, self.Kp = data, n_rows, Kp
        self.counts, self.row_off, self.row_item, self.mode = counts, row_off, row_item, mode


def pack_tokens(x, counts, *, slot0, mode=0, normalize=True, eps=1e-12, want_row_item=False, out=None,
                out_row_item=None, row_base=0, item_base=0):
    """x [B,S,d] fp32 cuda (any strides on B,S) -> Packed rows for tokens slot0 .. slot0+count-1."""
    lib = _cabi.lib()
    assert x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and x.stride(2) == 1
    B, S, d = x.shape
    counts = np.asarray(counts, dtype=np.int32)
    assert counts.shape == (B,)
    if B and (counts.min() < 0 or counts.max() > max(S - slot0, 0)):
        raise ValueError("token counts exceed the container")
    row_off, n_rows = exclusive_cumsum(counts)
    Kp = round_up(d * (1 if mode == 0 else 3), _cabi.TILE_K)
    if out is not None:
        # pack into rows [row_base, row_base + n_rows) of a caller-owned buffer (sharded packing)
        assert out.dtype == torch.bfloat16 and out.shape[1] == Kp and out.is_contiguous() and row_base + n_rows <= out.shape[0]
        data, row_item = out, out_row_item
        row_off = row_off + row_base
    else:
        data = torch.empty((max(n_rows, 1), Kp), dtype=torch.bfloat16, device=x.device)
        row_item = None
        if want_row_item:
            row_item = torch.full((max(padded_rows(n_rows), 2 * _cabi.TILE_M),), -1, dtype=torch.int32, device=x.device)
    if n_rows:
        cnt_d = _to_dev(counts, x.device)
        off_d = _to_dev(row_off, x.device)
        a = _cabi.PackArgs(
            src=x.data_ptr(), stride_b=x.stride(0), stride_s=x.stride(1), B=B, S=S, d=d, slot0=slot0,
            count=cnt_d.data_ptr(), row_off=off_d.data_ptr(), dst=data.data_ptr(), Kp=Kp, mode=mode,
            normalize=1 if normalize else 0, eps=eps, row_item=row_item.data_ptr() if row_item is not None else None,
            item_base=item_base)
        _cabi.check(lib.alad_pack_tokens(C.byref(a), _cabi.stream_ptr()), "alad_pack_tokens")
    return Packed(data, n_rows, Kp, counts, row_off, row_item, mode)


def mrsw_scores_packed(words, regions, tiles_dev, n_tiles, Ni, Nc, out=None, num_ctas=0, cta_group=0):
    """S[Ni,Nc] from packed operands (see alad_mrsw_scores_fwd)."""
    lib = _cabi.lib()
    assert words.Kp == regions.Kp
    dev = words.data.device
    if out is None:
        out = torch.empty((Ni, Nc), dtype=torch.float32, device=dev)
    assert out.shape == (Ni, Nc) and out.dtype == torch.float32 and out.stride(1) == 1
    a = _cabi.MrswFwdArgs(
        words=words.data.data_ptr(), n_word_rows=words.n_rows, regions=regions.data.data_ptr(),
        n_region_rows=regions.n_rows, Kp=words.Kp, row_cap=words.row_item.data_ptr(),
        ntiles=tiles_dev.data_ptr() if n_tiles else None, n_ntiles=n_tiles, S=out.data_ptr(),
        ldS=out.stride(0) if Ni > 1 else max(out.stride(0), Nc), Ni=Ni, Nc=Nc, epilogue=0, num_ctas=num_ctas,
        cta_group=cta_group, transpose_out=0)
    if kernel_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cabi.check(lib.alad_mrsw_scores_fwd(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_scores_fwd")
    if kernel_timeline is not None:
        e1.record()
        kernel_timeline.append((e0, e1, Ni, Nc, words.Kp))
    return out


def scored_counts(im_shape, s_shape, im_len, s_len):
    """(R, W, nr, nw, clamp): container extents, valid counts, 'has masked slots' flags."""
    R = max(im_shape[1] - 1, 0)
    W = max(s_shape[1] - 3, 0)
    nr = valid_counts(im_len, 1, R)
    nw = valid_counts(s_len, 3, W)
    if len(nr) != im_shape[0] or len(nw) != s_shape[0]:
        raise ValueError("length lists do not match the batch sizes")
    return R, W, nr, nw, nr < R


AGGREGATIONS = ("sum", "mean", "MrSw", "MrAVGw", "symm", "MwSr")


def _max_sum_scores(max_x, max_counts, max_clamp, sum_x, sum_counts, precision, transpose_out, out):
    """out[...] = sum over the valid tokens of `sum_x` items of the max over the valid tokens of
    `max_x` items (clamped at 0 where `max_clamp`).  The 'max' items become tile columns (N side),
    the 'sum' items packed rows (M side).  out is [n_max, n_sum], or [n_sum, n_max] if transpose_out."""
    split = precision == "fp32"
    rows = pack_tokens(sum_x, sum_counts, slot0=1, mode=1 if split else 0, want_row_item=True)
    cols = pack_tokens(max_x, max_counts, slot0=1, mode=2 if split else 0)
    _, table, _ = build_region_tiles(max_counts, max_clamp)
    tiles_dev = _to_dev(table.view(np.int32).reshape(-1), max_x.device) if len(table) else None
    n_max, n_sum = max_x.shape[0], sum_x.shape[0]
    lib = _cabi.lib()
    a = _cabi.MrswFwdArgs(
        words=rows.data.data_ptr(), n_word_rows=rows.n_rows, regions=cols.data.data_ptr(), n_region_rows=cols.n_rows,
        Kp=rows.Kp, row_cap=rows.row_item.data_ptr(), ntiles=tiles_dev.data_ptr() if len(table) else None,
        n_ntiles=len(table), S=out.data_ptr(), ldS=max(out.stride(0), out.shape[1]), Ni=n_max, Nc=n_sum, epilogue=0,
        num_ctas=0, cta_group=0, transpose_out=1 if transpose_out else 0)
    if kernel_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cabi.check(lib.alad_mrsw_scores_fwd(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_scores_fwd")
    if kernel_timeline is not None:
        e1.record()
        kernel_timeline.append((e0, e1, n_max, n_sum, rows.Kp))
    return out


def pool_tokens(x, counts, eps=1e-12):
    """[B,d] fp32: sum of the normalised valid tokens (slots 1..count) of every item."""
    B, S, d = x.shape
    out = torch.empty((B, d), dtype=torch.float32, device=x.device)
    if B:
        cnt = _to_dev(np.asarray(counts, np.int32), x.device)
        _cabi.check(_cabi.lib().alad_pool_tokens(x.data_ptr(), x.stride(0), x.stride(1), B, S, d, 1, cnt.data_ptr(), eps,
                                                 out.data_ptr(), _cabi.stream_ptr()), "alad_pool_tokens")
    return out


def scale_scores(S, mul=1.0, col_div=None):
    """In place S[i,j] = S[i,j] * mul / col_div[j]."""
    Ni, Nc = S.shape
    assert S.is_cuda and S.dtype == torch.float32 and S.stride(1) == 1
    _cabi.check(_cabi.lib().alad_scale_scores(S.data_ptr(), max(S.stride(0), Nc), Ni, Nc,
                                              col_div.data_ptr() if col_div is not None else None, float(mul),
                                              _cabi.stream_ptr()), "alad_scale_scores")
    return S


def unit_rows(x, eps=0.0):
    """Rows of x [B,d] scaled to unit L2 norm (alad.utils.l2norm, no eps by default): fp32 [B,d] on the
    device, produced by the pooling kernel with one token per item."""
    B, d = x.shape
    out = torch.empty((B, d), dtype=torch.float32, device=x.device)
    if B:
        cnt = torch.ones(B, dtype=torch.int32, device=x.device)
        _cabi.check(_cabi.lib().alad_pool_tokens(x.data_ptr(), x.stride(0), d, B, 1, d, 0, cnt.data_ptr(), eps,
                                                 out.data_ptr(), _cabi.stream_ptr()), "alad_pool_tokens")
    return out


def alignment_scores(im_set, s_seq, im_len, s_len, precision=None, out=None, aggregation="MrSw"):
    """Alignment scores S[B_i,B_c] of alad/loss.py:79-135 on the GPU, every tensor pooling mode:
    MrSw (max regions, sum words), MrAVGw (/ #words), MwSr (roles swapped), symm (MrSw + MwSr),
    sum / mean (one GEMM of the pooled token sums)."""
    if aggregation not in AGGREGATIONS:
        raise ValueError(f"unsupported aggregation {aggregation!r}")
    precision = precision or _precision
    im_set = _require_cuda(im_set, "im_set")
    s_seq = _require_cuda(s_seq, "s_seq")
    if im_set.dim() != 3 or s_seq.dim() != 3 or im_set.shape[2] != s_seq.shape[2]:
        raise ValueError("expected im_set [B_i,S_im,d] and s_seq [B_c,S_s,d] with equal d")
    Ni, Nc = im_set.shape[0], s_seq.shape[0]
    R, W, nr, nw, clamp = scored_counts(im_set.shape, s_seq.shape, im_len, s_len)
    if out is None:
        out = torch.empty((Ni, Nc), dtype=torch.float32, device=im_set.device)
    assert out.shape == (Ni, Nc) and out.dtype == torch.float32 and out.stride(1) == 1
    if aggregation in ("sum", "mean"):
        # sum_{r,w} <r,w> = <sum_r r, sum_w w>; always split precision (the GEMM is tiny)
        dot_scores(pool_tokens(im_set, nr), pool_tokens(s_seq, nw), precision="fp32", out=out)
        if aggregation == "mean":
            scale_scores(out, mul=1.0 / max(R * W, 1))
        return out
    if aggregation in ("MrSw", "MrAVGw", "symm"):
        _max_sum_scores(im_set, nr, clamp, s_seq, nw, precision, False, out)
        if aggregation == "MrAVGw":
            scale_scores(out, col_div=_to_dev(nw.astype(np.float32), out.device))
    if aggregation in ("MwSr", "symm"):
        dst = out if aggregation == "MwSr" else torch.empty_like(out)
        _max_sum_scores(s_seq, nw, nw < W, im_set, nr, precision, True, dst)
        if aggregation == "symm":
            out += dst
    return out


def dot_scores(im, s, precision=None, normalize=False, eps=0.0, out=None):
    """scores[B_i,B_c] = im @ s.T (alad/loss.py:8-11; cosine_sim :13-18 with normalize=True)
    through the same tcgen05 mainloop with the plain-GEMM epilogue."""
    lib = _cabi.lib()
    precision = precision or _precision
    im = _require_cuda(im, "im")
    s = _require_cuda(s, "s")
    if im.dim() != 2 or s.dim() != 2 or im.shape[1] != s.shape[1]:
        raise ValueError("expected im [B_i,d] and s [B_c,d]")
    Ni, Nc = im.shape[0], s.shape[0]
    split = precision == "fp32"
    words = pack_tokens(s.unsqueeze(1), np.ones(Nc, np.int32), slot0=0, mode=1 if split else 0,
                        normalize=normalize, eps=eps)
    regions = pack_tokens(im.unsqueeze(1), np.ones(Ni, np.int32), slot0=0, mode=2 if split else 0,
                          normalize=normalize, eps=eps)
    if out is None:
        out = torch.empty((Ni, Nc), dtype=torch.float32, device=im.device)
    if Ni == 0 or Nc == 0:
        return out
    table = gemm_tiles(Ni)
    tiles_dev = _to_dev(table.view(np.int32).reshape(-1), im.device)
    a = _cabi.MrswFwdArgs(
        words=words.data.data_ptr(), n_word_rows=Nc, regions=regions.data.data_ptr(), n_region_rows=Ni,
        Kp=words.Kp, row_cap=None, ntiles=tiles_dev.data_ptr(), n_ntiles=len(table), S=out.data_ptr(),
        ldS=max(out.stride(0), Nc), Ni=Ni, Nc=Nc, epilogue=1, num_ctas=0, cta_group=0, transpose_out=0)
    _cabi.check(lib.alad_mrsw_scores_fwd(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_scores_fwd")
    return out
