"""Drop-ins for the criteria of ``alad/loss.py`` used by ``ALADModel`` (alad/alad_model.py:275-292,
380, 386, 405): ``Contrastive``, ``AlignmentContrastiveLoss``, ``ContrastiveLoss``,
``DistillationLoss`` -- same constructor arguments, same ``forward`` signatures and return
conventions, no parameters (except ``DistillationLoss(mode='mse').wb``, kept so that
state_dict keys stay identical).  All math runs in the CUDA kernels behind the C ABI;
autograd is wired with ``torch.autograd.Function``."""
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _cabi, scan, scoring


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------
# kernels as functions
# --------------------------------------------------------------------------------------
def _rowmajor(x):
    """fp32, unit innermost stride and a row stride that really separates the rows (an expanded / broadcast matrix
    has stride(0) = 0: the kernels would read B*B floats from a buffer that does not hold them)."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() == 2 and x.shape[0] > 1 and (x.stride(1) != 1 or x.stride(0) < x.shape[1]):
        x = x.contiguous()
    elif x.dim() == 2 and x.stride(1) != 1:
        x = x.contiguous()
    return x


def triplet_fwd_bwd(scores, margin, max_violation, want_grad=True):
    """(loss 0-d, G [B,B] or None, row_arg, col_arg) -- alad/loss.py:42-67 + SURVEY A.1."""
    lib = _cabi.lib()
    if scores.dim() != 2 or scores.shape[0] != scores.shape[1]:
        raise RuntimeError("compute_contrastive_loss needs a square score matrix (torch.eye at alad/loss.py:55)")
    S = _rowmajor(scores.detach())
    B = S.shape[0]
    dev = S.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    G = torch.empty((B, B), dtype=torch.float32, device=dev) if want_grad else None
    ra = torch.empty(B, dtype=torch.int32, device=dev)
    ca = torch.empty(B, dtype=torch.int32, device=dev)
    ws = _ws(lib.alad_loss_workspace_bytes(B), dev)
    _cabi.check(lib.alad_triplet_fwd_bwd(S.data_ptr(), max(S.stride(0), B), B, float(margin), 1 if max_violation else 0,
                                         loss.data_ptr(), G.data_ptr() if G is not None else None, B, ra.data_ptr(),
                                         ca.data_ptr(), ws.data_ptr(), _cabi.stream_ptr()), "alad_triplet_fwd_bwd")
    return loss, G, ra, ca


def listnet_fwd_bwd(teacher, student, temperature=6.0, eps=1e-10, want_grad=True):
    """(loss 0-d, dM or None) -- alad/loss.py:427-445 + SURVEY A.2."""
    lib = _cabi.lib()
    T = teacher.detach()
    M = student.detach()
    if T.shape != M.shape or T.dim() != 2 or T.shape[0] != T.shape[1]:
        raise RuntimeError("listnet distillation expects two square matrices of equal shape")
    T, M = _rowmajor(T), _rowmajor(M)
    B = T.shape[0]
    dev = M.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dM = torch.empty((B, B), dtype=torch.float32, device=dev) if want_grad else None
    ws = _ws(lib.alad_loss_workspace_bytes(B), dev)
    _cabi.check(lib.alad_listnet_fwd_bwd(T.data_ptr(), max(T.stride(0), B), M.data_ptr(), max(M.stride(0), B), B,
                                         float(temperature), float(eps), loss.data_ptr(),
                                         dM.data_ptr() if dM is not None else None, B, ws.data_ptr(),
                                         _cabi.stream_ptr()), "alad_listnet_fwd_bwd")
    return loss, dM


def _square_pair(teacher, student, what):
    T = teacher.detach()
    M = student.detach()
    if T.shape != M.shape or T.dim() != 2 or T.shape[0] != T.shape[1]:
        raise RuntimeError(f"{what} distillation expects two square matrices of equal shape")
    return _rowmajor(T), _rowmajor(M)


def distill_mse_fwd_bwd(teacher, student, wb, want_grad=True):
    """(loss 0-d, dM or None, dwb [2] or None) -- alad/loss.py:371-373."""
    lib = _cabi.lib()
    T, M = _square_pair(teacher, student, "mse")
    B, dev = T.shape[0], M.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dM = torch.empty((B, B), dtype=torch.float32, device=dev) if want_grad else None
    dwb = torch.empty(2, dtype=torch.float32, device=dev) if want_grad else None
    wb_d = wb.detach().to(device=dev, dtype=torch.float32).contiguous()
    ws = _ws(lib.alad_distill_workspace_bytes(B, 0), dev)
    _cabi.check(lib.alad_distill_mse_fwd_bwd(T.data_ptr(), max(T.stride(0), B), M.data_ptr(), max(M.stride(0), B), B,
                                             wb_d.data_ptr(), loss.data_ptr(), dM.data_ptr() if want_grad else None, B,
                                             dwb.data_ptr() if want_grad else None, ws.data_ptr(), _cabi.stream_ptr()),
                "alad_distill_mse_fwd_bwd")
    return loss, dM, dwb


def distill_contrastive_fwd_bwd(teacher, student, margin, want_grad=True, mutate_teacher=True):
    """(loss 0-d, dM or None) -- alad/loss.py:397-418.  With mutate_teacher the diagonal of the caller's
    teacher matrix is zeroed in place, which is what the reference's `.detach().masked_fill_` does."""
    lib = _cabi.lib()
    M = _rowmajor(student.detach())
    T = teacher.detach()
    in_place = (mutate_teacher and T.is_cuda and T.dtype == torch.float32 and T.dim() == 2 and T.stride(1) == 1
                and (T.shape[0] <= 1 or T.stride(0) >= T.shape[1]))
    if not in_place:
        T = T.to(device=M.device, dtype=torch.float32).contiguous().clone()
    if T.shape != M.shape or T.dim() != 2 or T.shape[0] != T.shape[1]:
        raise RuntimeError("contrastive distillation expects two square matrices of equal shape")
    B, dev = T.shape[0], M.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dM = torch.empty((B, B), dtype=torch.float32, device=dev) if want_grad else None
    ws = _ws(lib.alad_distill_workspace_bytes(B, 1), dev)
    _cabi.check(lib.alad_distill_contrastive_fwd_bwd(T.data_ptr(), max(T.stride(0), B), M.data_ptr(), max(M.stride(0), B),
                                                     B, float(margin), 1 if mutate_teacher else 0, loss.data_ptr(),
                                                     dM.data_ptr() if want_grad else None, B, ws.data_ptr(),
                                                     _cabi.stream_ptr()), "alad_distill_contrastive_fwd_bwd")
    if mutate_teacher and not in_place:
        with torch.no_grad():
            teacher.detach().fill_diagonal_(0)       # same observable side effect for exotic layouts
    return loss, dM


def distill_ordinal_fwd_bwd(teacher, student, margin, threshold, stride, want_grad=True):
    """(loss 0-d, dM or None) -- alad/loss.py:374-396."""
    lib = _cabi.lib()
    T, M = _square_pair(teacher, student, "ordinal")
    B, dev = T.shape[0], M.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dM = torch.empty((B, B), dtype=torch.float32, device=dev) if want_grad else None
    ws = _ws(lib.alad_distill_workspace_bytes(B, 2), dev)
    _cabi.check(lib.alad_distill_ordinal_fwd_bwd(T.data_ptr(), max(T.stride(0), B), M.data_ptr(), max(M.stride(0), B), B,
                                                 float(margin), float(threshold), int(stride), loss.data_ptr(),
                                                 dM.data_ptr() if want_grad else None, B, ws.data_ptr(),
                                                 _cabi.stream_ptr()), "alad_distill_ordinal_fwd_bwd")
    return loss, dM


def normalize_bwd_(x, dx, eps):
    """In place dx <- J(x) dx for row-wise x / max(||x||, eps) (alad_normalize_bwd)."""
    rows, d = x.shape
    assert dx.shape == x.shape and x.stride(1) == 1 and dx.stride(1) == 1
    _cabi.check(_cabi.lib().alad_normalize_bwd(x.data_ptr(), x.stride(0), rows, d, float(eps), dx.data_ptr(), dx.stride(0),
                                               _cabi.stream_ptr()), "alad_normalize_bwd")
    return dx


def pool_tokens_bwd(x, counts, d_pool, eps=1e-12):
    """[B,S,d] gradient of scoring.pool_tokens (alad_pool_tokens_bwd)."""
    B, S, d = x.shape
    out = torch.empty((B, S, d), dtype=torch.float32, device=x.device)
    if B and S:
        cnt = scoring._to_dev(np.asarray(counts, np.int32), x.device)
        _cabi.check(_cabi.lib().alad_pool_tokens_bwd(x.data_ptr(), x.stride(0), x.stride(1), B, S, d, 1, cnt.data_ptr(),
                                                     float(eps), d_pool.data_ptr(), out.data_ptr(), _cabi.stream_ptr()),
                    "alad_pool_tokens_bwd")
    return out


def _grad_like(x):
    """Uninitialised fp32 [B,S,d] gradient buffer: the strides of x when x is a dense [S,B,d] tensor viewed as
    [B,S,d], contiguous otherwise."""
    B, S, d = x.shape
    if B > 1 and S > 1 and x.stride() == (d, B * d, 1):
        return torch.empty((S, B, d), dtype=torch.float32, device=x.device).permute(1, 0, 2)
    return torch.empty((B, S, d), dtype=torch.float32, device=x.device)


def mrsw_backward(im_set, s_seq, nr, nw, G0=None, g0_scale=None, G1=None, eps=1e-12, region_extent=0):
    """(d im_set, d s_seq) for dL/dS = g0_scale*G0 + G1 (alad_mrsw_scores_bwd)."""
    lib = _cabi.lib()
    Bi, S_im, d = im_set.shape
    Bc, S_s, _ = s_seq.shape
    dev = im_set.device
    # gradients are produced in the memory layout of the inputs: ALADModel hands over [S,B,d] tensors permuted to
    # [B,S,d] (alad_model.py:377-378), and a gradient with the same strides is accumulated into the leaf without a copy
    d_im = _grad_like(im_set)
    d_s = _grad_like(s_seq)
    max_pairs = max(Bi * Bc, 1)
    nbytes = lib.alad_mrsw_bwd_workspace_bytes(Bi, S_im, Bc, S_s, max_pairs)
    ws = _ws(nbytes, dev)
    nr_d, nw_d = scoring._to_dev_group([np.asarray(nr, np.int32), np.asarray(nw, np.int32)], dev)

    def ptr(t):
        return t.data_ptr() if t is not None else None

    for g in (G0, G1):
        assert g is None or (g.dtype == torch.float32 and g.stride(1) == 1 and g.shape == (Bi, Bc))
    a = _cabi.MrswBwdArgs(
        im=im_set.data_ptr(), im_stride_b=im_set.stride(0), im_stride_s=im_set.stride(1),
        s=s_seq.data_ptr(), s_stride_b=s_seq.stride(0), s_stride_s=s_seq.stride(1),
        Bi=Bi, S_im=S_im, Bc=Bc, S_s=S_s, d=d, nr=nr_d.data_ptr(), nw=nw_d.data_ptr(),
        G0=ptr(G0), ldG0=max(G0.stride(0), Bc) if G0 is not None else 0, g0_scale=ptr(g0_scale),
        G1=ptr(G1), ldG1=max(G1.stride(0), Bc) if G1 is not None else 0,
        d_im=d_im.data_ptr(), d_s=d_s.data_ptr(), eps=eps, region_extent=region_extent, max_pairs=max_pairs,
        workspace=ws.data_ptr(), workspace_bytes=nbytes,
        d_im_stride_b=d_im.stride(0), d_im_stride_s=d_im.stride(1), d_s_stride_b=d_s.stride(0), d_s_stride_s=d_s.stride(1))
    _cabi.check(lib.alad_mrsw_scores_bwd(C.byref(a), _cabi.stream_ptr()), "alad_mrsw_scores_bwd")
    return d_im, d_s


# --------------------------------------------------------------------------------------
# autograd wiring
# --------------------------------------------------------------------------------------
class _TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, margin, max_violation):
        loss, G, _, _ = triplet_fwd_bwd(scores, margin, max_violation, want_grad=scores.requires_grad)
        ctx.save_for_backward(G)
        return loss

    @staticmethod
    def backward(ctx, g):
        (G,) = ctx.saved_tensors
        return G * g, None, None


class _ListnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student):
        loss, dM = listnet_fwd_bwd(teacher, student, want_grad=student.requires_grad)
        ctx.save_for_backward(dM)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dM,) = ctx.saved_tensors
        return None, dM * g                     # teacher is detached (alad/loss.py:370)


class _DotScoresFn(torch.autograd.Function):
    """scores = im @ s.T on the tcgen05 GEMM; backward = two more GEMMs on the same kernel."""

    @staticmethod
    def forward(ctx, im, s, precision, normalize):
        ctx.save_for_backward(im, s)
        ctx.precision = precision
        ctx.normalize = normalize
        return scoring.dot_scores(im, s, precision=precision, normalize=normalize, eps=0.0)

    @staticmethod
    def backward(ctx, G):
        im, s = ctx.saved_tensors
        G = G.contiguous().float()
        im_f, s_f = im.detach().float().contiguous(), s.detach().float().contiguous()
        d_im = d_s = None
        # cosine_sim: the GEMM operands are the unit rows; the l2norm Jacobian (alad/utils.py:134-139,
        # no eps) is applied in place afterwards
        if ctx.needs_input_grad[0]:
            # G @ s(hat): "regions" = G rows [Bi, Bc], "words" = s columns [d, Bc]
            s_op = scoring.unit_rows(s_f) if ctx.normalize else s_f
            d_im = scoring.dot_scores(G, s_op.t().contiguous(), precision="fp32")
            if ctx.normalize:
                normalize_bwd_(im_f, d_im, 0.0)
        if ctx.needs_input_grad[1]:
            im_op = scoring.unit_rows(im_f) if ctx.normalize else im_f
            d_s = scoring.dot_scores(G.t().contiguous(), im_op.t().contiguous(), precision="fp32")   # G.T @ im(hat)
            if ctx.normalize:
                normalize_bwd_(s_f, d_s, 0.0)
        return d_im, d_s, None, None


class _OrderScoresFn(torch.autograd.Function):
    """order_sim (alad/loss.py:20-26) forward + gradient on CUDA cores (not a bilinear form)."""

    @staticmethod
    def forward(ctx, im, s):
        im_c = scoring._require_cuda(im.detach(), "im").contiguous()
        s_c = scoring._require_cuda(s.detach(), "s").contiguous()
        if im_c.dim() != 2 or s_c.dim() != 2 or im_c.shape[1] != s_c.shape[1]:
            raise ValueError("expected im [B_i,d] and s [B_c,d]")
        Ni, Nc, d = im_c.shape[0], s_c.shape[0], im_c.shape[1]
        out = torch.empty((Ni, Nc), dtype=torch.float32, device=im_c.device)
        _cabi.check(_cabi.lib().alad_order_scores(im_c.data_ptr(), d, s_c.data_ptr(), d, Ni, Nc, d, out.data_ptr(),
                                                  max(Nc, 1), _cabi.stream_ptr()), "alad_order_scores")
        ctx.save_for_backward(im_c, s_c, out)
        ctx.devices = (im.device, s.device)
        return out

    @staticmethod
    def backward(ctx, G):
        im_c, s_c, out = ctx.saved_tensors
        Ni, Nc, d = im_c.shape[0], s_c.shape[0], im_c.shape[1]
        G = G.contiguous().float()
        d_im = torch.empty_like(im_c)
        d_s = torch.empty_like(s_c)
        _cabi.check(_cabi.lib().alad_order_scores_bwd(im_c.data_ptr(), d, s_c.data_ptr(), d, Ni, Nc, d, out.data_ptr(),
                                                      max(Nc, 1), G.data_ptr(), max(Nc, 1), d_im.data_ptr(),
                                                      d_s.data_ptr(), _cabi.stream_ptr()), "alad_order_scores_bwd")
        return d_im.to(ctx.devices[0]), d_s.to(ctx.devices[1])


class _DistillMseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, wb):
        loss, dM, dwb = distill_mse_fwd_bwd(teacher, student, wb, want_grad=student.requires_grad or wb.requires_grad)
        ctx.save_for_backward(dM, dwb)
        ctx.wb_device = wb.device
        return loss

    @staticmethod
    def backward(ctx, g):
        dM, dwb = ctx.saved_tensors
        return None, dM * g, (dwb * g).to(ctx.wb_device)


class _DistillContrastiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, margin):
        loss, dM = distill_contrastive_fwd_bwd(teacher, student, margin, want_grad=student.requires_grad)
        ctx.save_for_backward(dM)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dM,) = ctx.saved_tensors
        return None, dM * g, None


class _DistillOrdinalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, margin, threshold, stride):
        loss, dM = distill_ordinal_fwd_bwd(teacher, student, margin, threshold, stride, want_grad=student.requires_grad)
        ctx.save_for_backward(dM)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dM,) = ctx.saved_tensors
        return None, dM * g, None, None, None


def alignment_backward(im_c, s_c, nr, nw, W, agg, G1=None, G0=None, g0_scale=None):
    """(d im_set, d s_seq) of the alignment scores for dL/dS = g0_scale * G0 + G1, every pooling mode
    of alad/loss.py:120-135 (im_c / s_c: raw fp32 CUDA tokens; nr / nw: valid scored counts)."""
    if agg == "MrSw":
        return mrsw_backward(im_c, s_c, nr, nw, G0=G0, g0_scale=g0_scale, G1=G1)
    G = G0 * g0_scale if G0 is not None and g0_scale is not None else G0
    if G1 is not None:
        G = G1 if G is None else G + G1
    if agg in ("sum", "mean"):
        # S = <pool(im), pool(s)> (/ (R*W)): two small GEMMs + the pooled-token Jacobian kernel
        if agg == "mean":
            G = G / max((im_c.shape[1] - 1) * W, 1)
        G = G.contiguous()
        pi, ps = scoring.pool_tokens(im_c, nr), scoring.pool_tokens(s_c, nw)
        d_pi = scoring.dot_scores(G, ps.t().contiguous(), precision="fp32")
        d_ps = scoring.dot_scores(G.t().contiguous(), pi.t().contiguous(), precision="fp32")
        return pool_tokens_bwd(im_c, nr, d_pi), pool_tokens_bwd(s_c, nw, d_ps)
    d_im = d_s = None
    if agg in ("MrAVGw", "symm"):
        Gm = G
        if agg == "MrAVGw":
            Gm = G / scoring._to_dev(np.asarray(nw, np.float32), G.device)[None, :]
            Gm = torch.nan_to_num(Gm, nan=0.0, posinf=0.0, neginf=0.0)
        d_im, d_s = mrsw_backward(im_c, s_c, nr, nw, G1=Gm.contiguous())
    if agg in ("MwSr", "symm"):
        # roles swapped: max over words (clamped when nw < W), sum over regions
        d_s2, d_im2 = mrsw_backward(s_c, im_c, nw, nr, G1=G.t().contiguous(), region_extent=max(W, 1))
        d_im = d_im2 if d_im is None else d_im + d_im2
        d_s = d_s2 if d_s is None else d_s + d_s2
    if d_im is None:
        raise ValueError(f"unsupported aggregation {agg!r}")
    return d_im, d_s


class _AlignmentFn(torch.autograd.Function):
    """(loss, S) = alignment scores + hinge; backward recomputes only the pairs whose dL/dS != 0."""

    @staticmethod
    def forward(ctx, im_set, s_seq, im_len, s_len, margin, max_violation, want_loss, precision, aggregation):
        ctx.set_materialize_grads(False)
        im_c = scoring._require_cuda(im_set.detach(), "im_set")
        s_c = scoring._require_cuda(s_seq.detach(), "s_seq")
        counts = scoring.scored_counts(im_c.shape, s_c.shape, im_len, s_len)
        _, W, nr, nw, _ = counts
        S = scoring.alignment_scores(im_c, s_c, im_len, s_len, precision=precision, aggregation=aggregation, counts=counts)
        needs_grad = im_set.requires_grad or s_seq.requires_grad
        G0 = None
        if want_loss:
            loss, G0, _, _ = triplet_fwd_bwd(S, margin, max_violation, want_grad=needs_grad)
        else:
            loss = torch.zeros((), dtype=torch.float32, device=S.device)
        ctx.nr, ctx.nw, ctx.W, ctx.aggregation = nr, nw, W, aggregation
        ctx.devices = (im_set.device, s_seq.device)
        ctx.save_for_backward(im_c, s_c, G0)
        return loss, S

    @staticmethod
    def backward(ctx, g_loss, g_S):
        im_c, s_c, G0 = ctx.saved_tensors
        use_G0 = G0 is not None and g_loss is not None
        none = (None,) * 9
        if not use_G0 and g_S is None:
            return none
        g_scale = g_loss.detach().float().reshape(1).contiguous() if use_G0 else None
        G1 = g_S.detach().float().contiguous() if g_S is not None else None
        d_im, d_s = alignment_backward(im_c, s_c, ctx.nr, ctx.nw, ctx.W, ctx.aggregation, G1, G0=G0 if use_G0 else None,
                                       g0_scale=g_scale)
        return (d_im.to(ctx.devices[0]) if ctx.needs_input_grad[0] else None,
                d_s.to(ctx.devices[1]) if ctx.needs_input_grad[1] else None) + none[2:]


# --------------------------------------------------------------------------------------
# nn.Module drop-ins
# --------------------------------------------------------------------------------------
def dot_sim(im, s):
    """alad/loss.py:8-11."""
    return _DotScoresFn.apply(im, s, None, False)


def cosine_sim(im, s):
    """alad/loss.py:13-18 (l2norm without eps, then mm)."""
    return _DotScoresFn.apply(im, s, None, True)


def order_sim(im, s):
    """alad/loss.py:20-26: -||max(0, s - im)||_2 for all pairs."""
    return _OrderScoresFn.apply(im, s)


class Contrastive(nn.Module):
    """alad/loss.py:29-67."""

    def __init__(self, margin=0, measure=False, max_violation=False):
        super().__init__()
        self.margin = margin
        if measure == 'order':
            self.sim = order_sim
        elif measure == 'cosine':
            self.sim = cosine_sim
        elif measure == 'dot':
            self.sim = dot_sim
        self.max_violation = max_violation

    def compute_contrastive_loss(self, scores):
        if not scores.is_cuda:
            scores = scores.cuda()
        return _TripletFn.apply(scores.float(), self.margin, self.max_violation)


class AlignmentContrastiveLoss(Contrastive):
    """alad/loss.py:70-159.  'MrSw' (every shipped config), 'MrAVGw', 'MwSr', 'symm' run the fused
    tcgen05 kernel; 'sum' / 'mean' collapse to one GEMM of pooled token sums; 'scan-sentences'
    (loss.py:136-149) is a plain GEMM of the unit token rows plus the per-pair attention kernels of
    csrc/scan_pool.cu (aladin_b200/scan.py)."""

    SUPPORTED = scoring.AGGREGATIONS + ("scan-sentences",)

    def __init__(self, margin=0, measure=False, max_violation=False, aggregation='sum-max-sentences'):
        super().__init__(margin, measure, max_violation)
        self.aggregation = aggregation
        self.precision = None          # None -> aladin_b200.get_precision()

    def forward(self, im_set, s_seq, im_len, s_len, return_loss=True, return_similarity_mat=False):
        if self.aggregation not in self.SUPPORTED:
            # the reference leaves `aggr_similarity` unbound for an unknown mode (alad/loss.py:120-151)
            raise UnboundLocalError(f"aggregation {self.aggregation!r} is not one of " + ", ".join(self.SUPPORTED))
        out_dev = im_set.device
        if self.aggregation == "scan-sentences":
            S = scan.ScanScoresFn.apply(im_set, s_seq, list(im_len), list(s_len), self.precision)
            loss = self.compute_contrastive_loss(S) if return_loss else None
        else:
            loss, S = _AlignmentFn.apply(im_set, s_seq, list(im_len), list(s_len), self.margin, self.max_violation,
                                         bool(return_loss), self.precision, self.aggregation)
        if out_dev.type != "cuda":
            loss, S = (loss.to(out_dev) if loss is not None else None), S.to(out_dev)
        if return_loss and return_similarity_mat:
            return loss, S
        elif return_loss:
            return loss
        elif return_similarity_mat:
            return S


class ContrastiveLoss(Contrastive):
    """alad/loss.py:162-186."""

    def __init__(self, margin=0, measure=False, max_violation=False):
        super().__init__(margin, measure, max_violation)

    def forward(self, im, s, return_similarity_mat=False):
        scores = self.sim(im, s)
        loss = self.compute_contrastive_loss(scores)
        if return_similarity_mat:
            return loss, scores
        return loss


class DistillationLoss(nn.Module):
    """alad/loss.py:359-447; modes 'mse', 'ordinal', 'contrastive', 'listnet' (every shipped config uses 'listnet')."""

    def __init__(self, mode='mse', margin=0.2, threshold=0.1, stride=3):
        super().__init__()
        self.mode = mode
        self.margin = margin
        self.threshold = threshold
        self.stride = stride
        if mode == 'mse':
            self.wb = nn.Parameter(torch.FloatTensor([0.5, 0.5]), requires_grad=True)

    def forward(self, teacher_scores, student_scores):
        if not student_scores.is_cuda:
            student_scores = student_scores.cuda()
        teacher = teacher_scores.detach()                     # alad/loss.py:370
        if self.mode == 'mse':
            return _DistillMseFn.apply(teacher.to(student_scores.device), student_scores, self.wb)
        elif self.mode == 'ordinal':
            return _DistillOrdinalFn.apply(teacher.to(student_scores.device), student_scores, self.margin, self.threshold,
                                           self.stride)
        elif self.mode == 'contrastive':
            return _DistillContrastiveFn.apply(teacher, student_scores, self.margin)
        elif self.mode == 'listnet':
            return _ListnetFn.apply(teacher.to(student_scores.device), student_scores)
        # the reference falls through to `return loss` with loss unbound (alad/loss.py:447)
        raise UnboundLocalError(f"DistillationLoss: unknown mode {self.mode!r}")
