"""Aggregation 'scan-sentences' of ``AlignmentContrastiveLoss.forward`` (alad/loss.py:136-149).

The reference expands both token sets to B x B x R x W x d (loss.py:143-146).  Here a pair needs only
its R x W block of region x word cosines and the W x W Gram matrix of the caption's unit word rows
(see csrc/scan_pool.cu): the cosines of ALL pairs come from one plain GEMM of the unit token rows on the
tcgen05 kernel (``scoring.dot_scores``), in image chunks that bound the materialised block; the per-pair
pooling and its gradient run in the CUDA-core kernels behind ``alad_scan_*``; the gradient w.r.t. the
tokens is two more GEMMs on the same tensor-core kernel.

Deviation kept on purpose: in the reference the gradient of this mode is NaN for every scored word slot
as soon as one image of the batch has a masked region (softmax over an all ``-inf`` row, loss.py:139-140,
then 0 * NaN in the backward of ``masked_fill_``); here masked regions simply get no gradient, which is
the reference's own result whenever that is finite (tests/golden/scan_sentences.npz, full-length images)."""
import numpy as np
import torch

from . import _cabi, scoring

_CHUNK_BYTES = 1 << 30          # upper bound of one materialised cosine block (fp32)
SPARSE_FRACTION = 8             # pair-list backward when at most 1 / SPARSE_FRACTION of dL/dS is non-zero


def _unit_token_rows(x, extent):
    """x [B,S,d] -> (raw [B*extent, d] contiguous copy of slots 1..extent, unit rows of it)."""
    B, _, d = x.shape
    raw = x[:, 1:1 + extent, :].reshape(B * extent, d)
    if not raw.is_contiguous():
        raw = raw.contiguous()
    return raw, scoring.unit_rows(raw, eps=1e-12)


def _gram(yh, Bc, W, nw_d):
    K = torch.empty((Bc, W, W), dtype=torch.float32, device=yh.device)
    _cabi.check(_cabi.lib().alad_scan_gram(yh.data_ptr(), Bc, W, yh.shape[1], nw_d.data_ptr(), K.data_ptr(),
                                           _cabi.stream_ptr()), "alad_scan_gram")
    return K


def _image_chunks(Bi, R, Bc, W):
    per_image = max(R * Bc * W * 4, 1)
    step = int(max(1, min(Bi, _CHUNK_BYTES // per_image)))
    return [(i0, min(i0 + step, Bi)) for i0 in range(0, Bi, step)]


def _prepare(im_c, s_c, counts):
    R, W, nr, nw, _ = counts
    if W == 0:
        # `im_len_mask[:, :, :, 0]` at alad/loss.py:147
        raise IndexError("index 0 is out of bounds for dimension 3 with size 0")
    dev = im_c.device
    nr_d, nw_d = scoring._to_dev_group([np.asarray(nr, np.int32), np.asarray(nw, np.int32)], dev)
    return R, W, nr_d, nw_d, int(nr.max()) if len(nr) else 0, int(nw.max()) if len(nw) else 0


def scan_scores(im_c, s_c, counts, precision=None, out=None):
    """S[B_i,B_c] of aggregation 'scan-sentences' (im_c / s_c: raw fp32 CUDA tokens)."""
    lib = _cabi.lib()
    Bi, Bc = im_c.shape[0], s_c.shape[0]
    R, W, nr_d, nw_d, max_nr, max_nw = _prepare(im_c, s_c, counts)
    if out is None:
        out = torch.empty((Bi, Bc), dtype=torch.float32, device=im_c.device)
    if Bi == 0 or Bc == 0:
        return out
    if R == 0:
        return out.zero_()
    _, xh = _unit_token_rows(im_c, R)
    _, yh = _unit_token_rows(s_c, W)
    K = _gram(yh, Bc, W, nw_d)
    for i0, i1 in _image_chunks(Bi, R, Bc, W):
        Cm = scoring.dot_scores(xh[i0 * R:i1 * R], yh, precision=precision)
        _cabi.check(lib.alad_scan_pool_fwd(Cm.data_ptr(), Cm.stride(0), i1 - i0, R, Bc, W, nr_d[i0:i1].data_ptr(),
                                           nw_d.data_ptr(), max_nr, max_nw, K.data_ptr(), out[i0:i1].data_ptr(),
                                           max(out.stride(0), Bc), _cabi.stream_ptr()), "alad_scan_pool_fwd")
    return out


def scan_backward(im_c, s_c, counts, G, precision=None):
    """(d im_set [B_i,S_im,d], d s_seq [B_c,S_s,d]) for dL/dS = G."""
    lib = _cabi.lib()
    Bi, S_im, d = im_c.shape
    Bc, S_s, _ = s_c.shape
    R, W, nr_d, nw_d, max_nr, max_nw = _prepare(im_c, s_c, counts)
    dev = im_c.device
    d_im = torch.zeros((Bi, S_im, d), dtype=torch.float32, device=dev)
    d_s = torch.zeros((Bc, S_s, d), dtype=torch.float32, device=dev)
    if Bi == 0 or Bc == 0 or R == 0 or max_nr == 0 or max_nw == 0:
        return d_im, d_s
    G = G.detach().float().contiguous()
    x_raw, xh = _unit_token_rows(im_c, R)
    y_raw, yh = _unit_token_rows(s_c, W)
    K = _gram(yh, Bc, W, nw_d)
    dK = torch.zeros_like(K)
    d_xh = torch.empty_like(xh)
    d_yh = torch.zeros_like(yh)
    # dL/dS with few non-zero entries (the hardest-negative hinge leaves <= 3B): the products with dC run
    # over the listed pairs only; otherwise two fp32-grade GEMMs on the tcgen05 kernel
    nz = torch.nonzero(G)
    sparse = nz.shape[0] * SPARSE_FRACTION <= Bi * Bc
    if sparse:
        d_xh.zero_()
    else:
        yh_t = yh.t().contiguous()
    for i0, i1 in _image_chunks(Bi, R, Bc, W):
        xh_c = xh[i0 * R:i1 * R]
        Cm = scoring.dot_scores(xh_c, yh, precision=precision)
        dC = torch.empty_like(Cm)
        _cabi.check(lib.alad_scan_pool_bwd(Cm.data_ptr(), Cm.stride(0), i1 - i0, R, Bc, W, nr_d[i0:i1].data_ptr(),
                                           nw_d.data_ptr(), max_nr, max_nw, K.data_ptr(), G[i0:i1].data_ptr(), G.stride(0),
                                           dC.data_ptr(), dC.stride(0), dK.data_ptr(), _cabi.stream_ptr()),
                    "alad_scan_pool_bwd")
        del Cm
        if sparse:
            sel = nz[(nz[:, 0] >= i0) & (nz[:, 0] < i1)]
            pairs = torch.stack([sel[:, 0] - i0, sel[:, 1]], dim=1).to(torch.int32).contiguous()
            _cabi.check(lib.alad_scan_apply_pairs(dC.data_ptr(), dC.stride(0), xh_c.data_ptr(), yh.data_ptr(),
                                                  pairs.data_ptr(), pairs.shape[0], i1 - i0, R, Bc, W, d,
                                                  nr_d[i0:i1].data_ptr(), nw_d.data_ptr(), max_nr, max_nw,
                                                  d_xh[i0 * R:i1 * R].data_ptr(), d_yh.data_ptr(), _cabi.stream_ptr()),
                        "alad_scan_apply_pairs")
        else:
            # d xhat = dC @ yhat, d yhat += dC.T @ xhat
            scoring.dot_scores(dC, yh_t, precision="fp32", out=d_xh[i0 * R:i1 * R])
            d_yh += scoring.dot_scores(dC.t().contiguous(), xh_c.t().contiguous(), precision="fp32")
        del dC
    _cabi.check(lib.alad_scan_gram_bwd(yh.data_ptr(), Bc, W, d, nw_d.data_ptr(), dK.data_ptr(), d_yh.data_ptr(),
                                       _cabi.stream_ptr()), "alad_scan_gram_bwd")
    # Jacobian of F.normalize (alad/loss.py:80-81) in place, then back into the dropped-slot layout
    _normalize_bwd_(x_raw, d_xh)
    _normalize_bwd_(y_raw, d_yh)
    d_im[:, 1:1 + R, :] = d_xh.view(Bi, R, d)
    d_s[:, 1:1 + W, :] = d_yh.view(Bc, W, d)
    return d_im, d_s


def _normalize_bwd_(x, dx):
    rows, d = x.shape
    _cabi.check(_cabi.lib().alad_normalize_bwd(x.data_ptr(), x.stride(0), rows, d, 1e-12, dx.data_ptr(), dx.stride(0),
                                               _cabi.stream_ptr()), "alad_normalize_bwd")


class ScanScoresFn(torch.autograd.Function):
    """S = scan-sentences scores; the hinge on top of it is a separate autograd node (loss._TripletFn)."""

    @staticmethod
    def forward(ctx, im_set, s_seq, im_len, s_len, precision):
        im_c = scoring._require_cuda(im_set.detach(), "im_set")
        s_c = scoring._require_cuda(s_seq.detach(), "s_seq")
        if im_c.dim() != 3 or s_c.dim() != 3 or im_c.shape[2] != s_c.shape[2]:
            raise ValueError("expected im_set [B_i,S_im,d] and s_seq [B_c,S_s,d] with equal d")
        counts = scoring.scored_counts(im_c.shape, s_c.shape, im_len, s_len)
        ctx.counts, ctx.precision = counts, precision
        ctx.devices = (im_set.device, s_seq.device)
        ctx.save_for_backward(im_c, s_c)
        return scan_scores(im_c, s_c, counts, precision=precision)

    @staticmethod
    def backward(ctx, G):
        im_c, s_c = ctx.saved_tensors
        d_im, d_s = scan_backward(im_c, s_c, ctx.counts, G, precision=ctx.precision)
        return (d_im.to(ctx.devices[0]) if ctx.needs_input_grad[0] else None,
                d_s.to(ctx.devices[1]) if ctx.needs_input_grad[1] else None, None, None, None)
