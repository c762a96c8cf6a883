"""Device-side scoring ops (thin wrappers over the C ABI): token packing, the fused
MrSw alignment scores and the global-vector dot scores.

Reference call sites replaced (mesnico/ALADIN):
  alad/loss.py:79-125  AlignmentContrastiveLoss.forward (aggregation 'MrSw')
  alad/loss.py:8-18    dot_sim / cosine_sim
"""
import collections
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .tiling import exclusive_cumsum, padded_rows, round_up, valid_counts

PRECISIONS = ("bf16", "fp32", "tf32")
# per precision code (0 bf16, 1 split-precision fp32, 2 tf32): pack mode of the word / region rows, row width in units of d
PRECISION_CODE = {"bf16": 0, "fp32": 1, "tf32": 2}
WORD_MODE = (0, 1, 3)
REGION_MODE = (0, 2, 3)
K_FACTOR = (1, 3, 2)
# bench.py sets this to a list to collect (start_event, end_event, pairs, Kp) of every scoring launch
kernel_timeline = None
_precision = "bf16"


def set_precision(mode):
    """'bf16': bf16 operands, fp32 accumulate (<= 1e-2 abs on scores).
    'tf32': fp32 operands rounded to TF32 on the tensor cores' TF32 path (retrieval galleries: i2t / t2i /
            AlignmentGallery): worst score entry 3e-4 .. 6e-4 relative -- NOT the 1e-4 parity mode, that is 'fp32' -- at half
            the time of 'fp32' (COCO-5k: 0.74 s against 1.5 s; 'bf16' 0.33 s).  The small training-batch calls take the
            'fp32' path for it.
    'fp32': split-precision hi/lo bf16 operands (3 tensor-core products per dot product,
    fp32-grade results, <= 1e-4 rel)."""
    global _precision
    if mode not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}")
    _precision = mode


def get_precision():
    return _precision


_cuda_checked = False


def _cuda_ok():
    global _cuda_checked
    if not _cuda_checked:
        _cuda_checked = torch.cuda.is_available()      # only a positive answer is remembered
    return _cuda_checked


def _require_cuda(x, name):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not _cuda_ok():
        raise _cabi.AladError("aladin_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    if x.device.type != "cuda":
        x = x.cuda(non_blocking=True)
    elif x.device.index is not None and x.device.index != torch.cuda.current_device():
        # the C ABI launches on the CURRENT device's stream (include/alad_b200.h): pointers of another device would be
        # dereferenced there
        raise _cabi.AladError(f"{name} lives on cuda:{x.device.index} but the current device is cuda:{torch.cuda.current_device()}: "
                              f"call torch.cuda.set_device({x.device.index}) or wrap the call in torch.cuda.device(...)")
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() >= 1 and x.stride(-1) != 1:
        x = x.contiguous()
    return x


# Small per-call metadata (token counts, packed row offsets, tile tables) is derived on the host from the Python
# length lists the reference passes around (alad/dataset.py:358-361), so every call needs it on the device.  Uploads
# are memoised by content, device and stream: the shape-only tables of dot_scores, and the lengths a backward pass
# shares with its forward, hit the cache; a new batch of lengths costs one merged upload per operand group.
_META_CACHE = collections.OrderedDict()
_META_CACHE_ENTRIES = 256
_META_CACHE_MAX_BYTES = 1 << 16
_TORCH_DTYPE = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.uint32): torch.int32,
                np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8}


def _to_dev(arr, device):
    arr = np.ascontiguousarray(arr)
    if arr.nbytes > _META_CACHE_MAX_BYTES:
        return torch.from_numpy(arr).to(device, non_blocking=True)
    device = torch.device(device)
    index = device.index if device.index is not None else torch._C._cuda_getDevice()
    key = (index, torch._C._cuda_getCurrentRawStream(index), arr.dtype.num, arr.shape, arr.tobytes())
    hit = _META_CACHE.get(key)
    if hit is not None:
        _META_CACHE.move_to_end(key)
        return hit
    if arr.dtype == np.uint32:
        arr = arr.view(np.int32)
    out = torch.from_numpy(arr).to(device, non_blocking=True)
    _META_CACHE[key] = out
    if len(_META_CACHE) > _META_CACHE_ENTRIES:
        _META_CACHE.popitem(last=False)
    return out


def _to_dev_group(arrays, device):
    """Several small host arrays -> ONE upload; returns typed device views (16-byte aligned)."""
    arrs = [np.ascontiguousarray(a) for a in arrays]
    offs, total = [], 0
    for a in arrs:
        offs.append(total)
        total += (a.nbytes + 15) & ~15
    blob = np.zeros(max(total, 16), dtype=np.uint8)
    for a, o in zip(arrs, offs):
        blob[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
    d = _to_dev(blob, device)
    return [d[o:o + a.nbytes].view(_TORCH_DTYPE[a.dtype]).reshape(a.shape) for a, o in zip(arrs, offs)]


class Packed:
    """Dense K-major bf16 token rows on the device."""
    __slots__ = ("data", "n_rows", "Kp", "counts", "row_off", "row_item", "mode")

    def __init__(self, data, n_rows, Kp, counts, row_off, row_item, mode):
        self.data, self.n_rows, self.Kp = data, n_rows, Kp
        self.counts, self.row_off, self.row_item, self.mode = counts, row_off, row_item, mode


def pack_meta(counts, row_base=0):
    """Host arrays (counts int32, row_off int64, n_rows) of one pack call."""
    counts = np.asarray(counts, dtype=np.int32)
    row_off, n_rows = exclusive_cumsum(counts)
    return counts, row_off + row_base if row_base else row_off, n_rows


def pack_tokens(x, counts, *, slot0, mode=0, normalize=True, eps=1e-12, want_row_item=False, out=None,
                out_row_item=None, row_base=0, item_base=0, meta_dev=None):
    """x [B,S,d] fp32 cuda (any strides on B,S) -> Packed rows for tokens slot0 .. slot0+count-1."""
    lib = _cabi.lib()
    assert x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and x.stride(2) == 1
    B, S, d = x.shape
    counts = np.asarray(counts, dtype=np.int32)
    assert counts.shape == (B,)
    if meta_dev is not None and len(meta_dev) == 4:
        row_off, n_rows = meta_dev[2], meta_dev[3]       # host offsets / row count computed by the caller
    else:
        if B and (counts.min() < 0 or counts.max() > max(S - slot0, 0)):
            raise ValueError("token counts exceed the container")
        row_off, n_rows = exclusive_cumsum(counts)
    Kp = round_up(d * (1 if mode == 0 else 2 if mode == 3 else 3), _cabi.TILE_K)      # in 2-byte units (mode 3: fp32 rows)
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
        # meta_dev = (counts, row offsets incl. row_base) already on the device (merged upload by the caller)
        off_d, cnt_d = meta_dev[:2] if meta_dev is not None else _to_dev_group([row_off, counts], x.device)
        a = _cabi.PackArgs(
            src=x.data_ptr(), stride_b=x.stride(0), stride_s=x.stride(1), B=B, S=S, d=d, slot0=slot0,
            count=cnt_d.data_ptr(), row_off=off_d.data_ptr(), dst=data.data_ptr(), Kp=Kp, mode=mode,
            normalize=1 if normalize else 0, eps=eps, row_item=row_item.data_ptr() if row_item is not None else None,
            item_base=item_base)
        _cabi.check(lib.alad_pack_tokens(C.byref(a), _cabi.stream_ptr()), "alad_pack_tokens")
    return Packed(data, n_rows, Kp, counts, row_off, row_item, mode)


def mrsw_scores_packed(words, regions, tiles_dev, n_tiles, Ni, Nc, out=None, num_ctas=0, cta_group=0, accumulate=False,
                       out_ptr=None, timeline_nc=None):
    """S[Ni,Nc] from packed operands (see alad_mrsw_scores_fwd).  accumulate: add onto a matrix the caller zeroed;
    out_ptr: raw device address of a dense [Ni, Nc] fp32 matrix instead of `out` (another GPU's block through a
    peer mapping)."""
    lib = _cabi.lib()
    assert words.Kp == regions.Kp and (words.mode == 3) == (regions.mode == 3), "operands packed for different precisions"
    dev = words.data.device
    if out_ptr is None:
        if out is None:
            out = torch.empty((Ni, Nc), dtype=torch.float32, device=dev)
        assert out.shape == (Ni, Nc) and out.dtype == torch.float32 and out.stride(1) == 1
        s_ptr, s_ld = out.data_ptr(), (out.stride(0) if Ni > 1 else max(out.stride(0), Nc))
    else:
        s_ptr, s_ld = out_ptr, Nc
    a = _cabi.MrswFwdArgs(
        words=words.data.data_ptr(), n_word_rows=words.n_rows, regions=regions.data.data_ptr(),
        n_region_rows=regions.n_rows, Kp=words.Kp, row_cap=words.row_item.data_ptr(),
        ntiles=tiles_dev.data_ptr() if n_tiles else None, n_ntiles=n_tiles, S=s_ptr,
        ldS=s_ld, Ni=Ni, Nc=Nc, epilogue=0, num_ctas=num_ctas,
        cta_group=cta_group, transpose_out=0, accumulate=1 if accumulate else 0,
        operand_format=1 if words.mode == 3 else 0)
    if kernel_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cabi.check(lib.alad_mrsw_scores_fwd(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_scores_fwd")
    if kernel_timeline is not None:
        e1.record()
        kernel_timeline.append((e0, e1, Ni, Nc if timeline_nc is None else timeline_nc, words.Kp))   # timeline_nc: captions' worth of rows of a partial launch
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


def _scores_fused(max_x, max_counts, max_clamp, sum_x, sum_counts, slot0, precision, epilogue, normalize, eps,
                  transpose_out, out):
    """alad_scores_fused: pack both operands and score them in one native call (host bookkeeping included)."""
    lib = _cabi.lib()
    n_max, S_max, d = max_x.shape
    n_sum, S_sum, _ = sum_x.shape
    max_counts = np.ascontiguousarray(max_counts, dtype=np.int32)
    sum_counts = np.ascontiguousarray(sum_counts, dtype=np.int32)
    clamp = np.ascontiguousarray(max_clamp, dtype=np.uint8) if max_clamp is not None else None
    split = 1 if precision in ("fp32", "tf32") else 0      # the fused small-batch call has no TF32 variant: 'tf32' -> split precision
    nbytes = lib.alad_scores_fused_workspace_bytes(n_max, S_max, slot0, n_sum, S_sum, slot0, d, split)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=max_x.device)
    a = _cabi.ScoresFusedArgs(
        max_x.data_ptr(), max_x.stride(0), max_x.stride(1), sum_x.data_ptr(), sum_x.stride(0), sum_x.stride(1),
        n_max, S_max, slot0, n_sum, S_sum, slot0, d, max_counts.ctypes.data, sum_counts.ctypes.data,
        clamp.ctypes.data if clamp is not None else None, split, epilogue, 1 if normalize else 0, eps,
        out.data_ptr(), max(out.stride(0), out.shape[1]), 1 if transpose_out else 0, ws.data_ptr(), nbytes)
    if kernel_timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cabi.check(lib.alad_scores_fused(C.byref(a), _cabi.stream_ptr()), "alad_scores_fused")
    if kernel_timeline is not None:
        e1.record()
        kernel_timeline.append((e0, e1, n_max, n_sum, round_up(d * (3 if split else 1), _cabi.TILE_K)))
    return out


def _max_sum_scores(max_x, max_counts, max_clamp, sum_x, sum_counts, precision, transpose_out, out):
    """out[...] = sum over the valid tokens of `sum_x` items of the max over the valid tokens of
    `max_x` items (clamped at 0 where `max_clamp`).  The 'max' items become tile columns (N side),
    the 'sum' items packed rows (M side).  out is [n_max, n_sum], or [n_sum, n_max] if transpose_out."""
    if len(max_counts) and int(np.max(max_counts)) > _cabi.TILE_N:
        raise ValueError(f"an item has {int(np.max(max_counts))} scored tokens on the max side; the kernel supports at most "
                         f"{_cabi.TILE_N}")
    return _scores_fused(max_x, max_counts, max_clamp, sum_x, sum_counts, 1, precision, 0, True, 1e-12, transpose_out, out)


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


def alignment_scores(im_set, s_seq, im_len, s_len, precision=None, out=None, aggregation="MrSw", counts=None):
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
    R, W, nr, nw, clamp = counts if counts is not None else scored_counts(im_set.shape, s_seq.shape, im_len, s_len)
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


_ONES = {}           # n -> np.ones(n, int32): the "one row per item" count arrays of dot_scores


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
    if out is None:
        out = torch.empty((Ni, Nc), dtype=torch.float32, device=im.device)
    if Ni == 0 or Nc == 0:
        return out
    assert out.shape == (Ni, Nc) and out.dtype == torch.float32 and out.stride(1) == 1
    ones = _ONES.get(max(Ni, Nc))
    if ones is None:
        if len(_ONES) >= 64:
            _ONES.clear()
        ones = _ONES[max(Ni, Nc)] = np.ones(max(Ni, Nc), np.int32)
    _scores_fused(im.unsqueeze(1), ones[:Ni], None, s.unsqueeze(1), ones[:Nc], 0, precision, 1, normalize, eps, False, out)
    return out
