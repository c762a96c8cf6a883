"""Drop-in for the loss call site of the reference model: ``ALADModel.forward_loss``
(alad/alad_model.py:371-428), the method that invokes the three criteria of the hot path

    matching_criterion(img_emb, cap_emb, return_similarity_mat=True)                      # :380
    alignment_criterion(img_emb_set, cap_emb_seq, img_lengths, cap_lengths, ...)          # :386
    distillation_loss(teacher_scores, matching_mat)                                       # :405

and logs each loss with ``.item()`` (three host syncs).  With the per-criterion drop-ins of
``aladin_b200.loss`` a training step at B <= 512 is bound by the HOST: ~35 kernel launches of a few
microseconds each behind a dozen Python-level calls and two trips through the autograd engine per criterion.
``train_losses`` runs the same stack through ``alad_train_losses_fwd`` / ``alad_train_losses_bwd``: one native
call and one autograd node per direction, one device->host read for the logger.

Usage (reference side)::

    from aladin_b200 import alad_model as fused
    fused.install(ALADModel)          # ALADModel.forward_loss -> fused path when the configuration allows it

The fused path covers what every shipped config selects (configs/*.yaml: measure 'dot', alignment-mode 'MrSw',
distillation-mode 'listnet', loss-type subsets of alignment / matching / distillation).  Any other
configuration ('selfaggregation', 'entropy', another pooling or distillation mode, criteria that are not the
aladin_b200 drop-ins) is handed to the model's original ``forward_loss`` untouched."""
import ctypes as C

import numpy as np
import torch

from . import _cabi, loss as L, scoring

LOSS_NAMES = ("matching", "alignment", "distillation")


def _rowmajor2d(x):
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


class _TrainLossesFn(torch.autograd.Function):
    """(matching_loss, alignment_loss, distillation_loss, matching_mat, teacher_scores); the two matrices are
    not differentiable outputs (the reference only feeds them to the distillation loss, which is inside)."""

    @staticmethod
    def forward(ctx, img_emb, cap_emb, im_set, s_seq, im_len, s_len, margin_m, mv_m, margin_a, mv_a, with_distill,
                precision, precision_m):
        lib = _cabi.lib()
        ctx.set_materialize_grads(False)
        im_cls = _rowmajor2d(scoring._require_cuda(img_emb, "img_emb"))
        s_cls = _rowmajor2d(scoring._require_cuda(cap_emb, "cap_emb"))
        im_c = scoring._require_cuda(im_set.detach(), "im_set")
        s_c = scoring._require_cuda(s_seq.detach(), "s_seq")
        B, d = im_cls.shape
        if s_cls.shape != (B, d) or im_c.dim() != 3 or s_c.dim() != 3 or im_c.shape[0] != B or s_c.shape[0] != B \
                or im_c.shape[2] != d or s_c.shape[2] != d:
            raise ValueError("train_losses expects img_emb/cap_emb [B,d], im_set [B,S_im,d], s_seq [B,S_s,d]")
        _, _, nr, nw, clamp = scoring.scored_counts(im_c.shape, s_c.shape, im_len, s_len)
        if B and int(nr.max()) > _cabi.TILE_N:
            raise ValueError(f"an image has {int(nr.max())} scored regions; the kernel supports at most {_cabi.TILE_N}")
        nr = np.ascontiguousarray(nr, dtype=np.int32)
        nw = np.ascontiguousarray(nw, dtype=np.int32)
        clamp = np.ascontiguousarray(clamp, dtype=np.uint8)
        dev = im_cls.device
        needs = [bool(t.requires_grad) for t in (img_emb, cap_emb, im_set, s_seq)]
        want_grad = any(needs)
        # ('tf32' exists for the retrieval galleries only: the fused small-batch call takes the split-precision path for it)
        split = 1 if (precision or scoring.get_precision()) in ("fp32", "tf32") else 0
        split_m = 1 if (precision_m or scoring.get_precision()) in ("fp32", "tf32") else 0
        # one buffer: losses[3] | pad | M | S | G_m | G_a | dM
        bb = B * B
        n_mat = 5 if want_grad else 2
        buf = torch.empty(64 + n_mat * bb, dtype=torch.float32, device=dev)
        mats = [buf[64 + k * bb:64 + (k + 1) * bb].view(B, B) for k in range(n_mat)]
        nbytes = lib.alad_train_losses_workspace_bytes(B, im_c.shape[1], s_c.shape[1], d, split, 0)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        a = _cabi.TrainLossesArgs(
            im_cls=im_cls.data_ptr(), ld_im_cls=im_cls.stride(0), s_cls=s_cls.data_ptr(), ld_s_cls=s_cls.stride(0),
            im_set=im_c.data_ptr(), im_stride_b=im_c.stride(0), im_stride_s=im_c.stride(1),
            s_seq=s_c.data_ptr(), s_stride_b=s_c.stride(0), s_stride_s=s_c.stride(1),
            B=B, S_im=im_c.shape[1], S_s=s_c.shape[1], d=d,
            nr=nr.ctypes.data, nw=nw.ctypes.data, clamp=clamp.ctypes.data, precision=split, precision_m=split_m,
            margin_m=float(margin_m), max_violation_m=1 if mv_m else 0, margin_a=float(margin_a),
            max_violation_a=1 if mv_a else 0, with_distill=1 if with_distill else 0, temperature=6.0, listnet_eps=1e-10,
            want_grad=1 if want_grad else 0, losses=buf.data_ptr(), M=mats[0].data_ptr(), S=mats[1].data_ptr(),
            G_m=mats[2].data_ptr() if want_grad else None, G_a=mats[3].data_ptr() if want_grad else None,
            dM=mats[4].data_ptr() if want_grad else None, workspace=ws.data_ptr(), workspace_bytes=ws.numel())
        _cabi.check(lib.alad_train_losses_fwd(C.byref(a), _cabi.stream_ptr()), "alad_train_losses_fwd")
        ctx.args = a                       # shapes, strides and scalars are reused by the backward call
        ctx.keep = (im_cls, s_cls, im_c, s_c, buf, nr, nw, clamp)     # keeps the pointers inside `a` alive
        ctx.needs = needs
        ctx.devices = (img_emb.device, cap_emb.device, im_set.device, s_seq.device)
        ctx.want_grad = want_grad
        M, S = mats[0], mats[1]
        ctx.mark_non_differentiable(M, S)
        return buf[0], buf[1], buf[2], M, S

    @staticmethod
    def backward(ctx, g_m, g_a, g_d, _gM, _gS):
        none = (None,) * 13
        if not ctx.want_grad or (g_m is None and g_a is None and g_d is None):
            return none
        lib = _cabi.lib()
        a = ctx.args
        im_cls, s_cls, im_c, s_c, buf, nr, nw, clamp = ctx.keep
        dev = buf.device
        B, d = a.B, a.d
        zero = buf.new_zeros(())
        g = torch.stack([x.detach().float().reshape(()) if x is not None else zero for x in (g_m, g_a, g_d)])
        need_cls = (ctx.needs[0] or ctx.needs[1]) and (g_m is not None or g_d is not None)
        need_set = (ctx.needs[2] or ctx.needs[3]) and g_a is not None
        d_icls = torch.empty((B, d), dtype=torch.float32, device=dev) if need_cls and ctx.needs[0] else None
        d_ccls = torch.empty((B, d), dtype=torch.float32, device=dev) if need_cls and ctx.needs[1] else None
        d_im = L._grad_like(im_c) if need_set else None
        d_s = L._grad_like(s_c) if need_set else None
        nbytes = lib.alad_train_losses_workspace_bytes(B, a.S_im, a.S_s, d, a.precision, 1)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        a.g = g.data_ptr()
        a.has_g_m, a.has_g_a, a.has_g_d = int(g_m is not None), int(g_a is not None), int(g_d is not None)
        a.d_im_cls = d_icls.data_ptr() if d_icls is not None else None
        a.d_s_cls = d_ccls.data_ptr() if d_ccls is not None else None
        a.d_im_set = d_im.data_ptr() if need_set else None
        a.d_s_seq = d_s.data_ptr() if need_set else None
        if need_set:
            a.d_im_stride_b, a.d_im_stride_s = d_im.stride(0), d_im.stride(1)
            a.d_s_stride_b, a.d_s_stride_s = d_s.stride(0), d_s.stride(1)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _cabi.check(lib.alad_train_losses_bwd(C.byref(a), _cabi.stream_ptr()), "alad_train_losses_bwd")
        dv = ctx.devices
        return (d_icls.to(dv[0]) if d_icls is not None else None,
                d_ccls.to(dv[1]) if d_ccls is not None else None,
                d_im.to(dv[2]) if need_set and ctx.needs[2] else None,
                d_s.to(dv[3]) if need_set and ctx.needs[3] else None) + none[4:]


def train_losses(img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, margin=0.2, max_violation=True,
                 with_distillation=True, precision=None, margin_alignment=None, max_violation_alignment=None,
                 precision_matching=None):
    """The three losses of ALADModel.forward_loss for measure 'dot' / aggregation 'MrSw' / distillation 'listnet'.

    img_emb, cap_emb         [B, d]      matching-head vectors (alad_model.py:240-241)
    img_emb_set, cap_emb_seq [B, S, d]   token outputs, batch first (i.e. after the permute of :377-378; any strides)
    precision / precision_matching: 'bf16' | 'fp32' | None (= aladin_b200.get_precision()) for the alignment kernel and
    the matching GEMM (the per-criterion drop-ins take them from criterion.precision and the global mode).
    Returns (matching_loss, alignment_loss, distillation_loss, matching_mat [B,B], teacher_scores [B,B])."""
    ma = margin if margin_alignment is None else margin_alignment
    va = max_violation if max_violation_alignment is None else max_violation_alignment
    return _TrainLossesFn.apply(img_emb, cap_emb, img_emb_set, cap_emb_seq, list(img_lengths), list(cap_lengths),
                                margin, bool(max_violation), ma, bool(va), bool(with_distillation), precision,
                                precision_matching)


def fused_eligible(model):
    """True when `model` (an ALADModel) is configured the way every shipped config is and owns the aladin_b200
    criteria, so that forward_loss can run as one fused call."""
    types = set(getattr(model, "losses_types", ()) or ())
    if not types or not types <= set(LOSS_NAMES) or not ({"alignment", "distillation"} & types):
        return False
    mc, ac = getattr(model, "matching_criterion", None), getattr(model, "alignment_criterion", None)
    if type(mc) is not L.ContrastiveLoss or getattr(mc, "sim", None) is not L.dot_sim:
        return False
    if type(ac) is not L.AlignmentContrastiveLoss or ac.aggregation != "MrSw":
        return False
    if "distillation" in types:
        dl = getattr(model, "distillation_loss", None)
        if type(dl) is not L.DistillationLoss or dl.mode != "listnet":
            return False
    return True


def forward_loss(self, img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, reg_loss=None):
    """Fused body of ALADModel.forward_loss (alad/alad_model.py:371-428): same arguments ([S,B,d] token tensors,
    permuted here like :377-378), same `losses` dict (keys in the reference's insertion order), same logger keys --
    fed from ONE device->host read instead of three `.item()` syncs."""
    types = self.losses_types
    mc, ac = self.matching_criterion, self.alignment_criterion
    lm, la, ld, _, _ = train_losses(
        img_emb, cap_emb, img_emb_set.permute(1, 0, 2), cap_emb_seq.permute(1, 0, 2), img_lengths, cap_lengths,
        margin=mc.margin, max_violation=mc.max_violation, with_distillation="distillation" in types, precision=ac.precision,
        margin_alignment=ac.margin, max_violation_alignment=ac.max_violation)
    losses = {}
    logged = []
    if "matching" in self.config["training"]["loss-type"]:
        losses["matching"] = lm
        logged.append(("matching_loss", 0, img_emb.size(0)))
    if "alignment" in types:
        losses["alignment"] = la
        logged.append(("alignment_loss", 1, img_emb_set.size(1)))
    if "distillation" in types:
        losses["distillation"] = ld
        logged.append(("distillation_loss", 2, img_emb.size(0)))
    logger = getattr(self, "logger", None)
    if logger is not None and logged:
        vals = torch.stack([lm.detach(), la.detach(), ld.detach()]).tolist()      # one sync
        for key, k, n in logged:
            logger.update(key, vals[k], n)
    return losses


def install(model_cls):
    """Patch `model_cls.forward_loss` (the reference's ALADModel): fused path when `fused_eligible`, the original
    method otherwise.  Idempotent; returns the class."""
    original = model_cls.forward_loss
    if getattr(original, "_alad_b200_fused", False):
        return model_cls

    def patched(self, img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, reg_loss):
        if fused_eligible(self):
            return forward_loss(self, img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, reg_loss)
        return original(self, img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, reg_loss)

    patched._alad_b200_fused = True
    patched._alad_b200_original = original
    model_cls.forward_loss = patched
    return model_cls
