"""``torch.ops.alad_b200.*`` -- the thin torch custom-op layer over the C ABI.

Every op is a ``torch.library.custom_op`` whose implementation is one or two calls into
``libalad_b200.so`` (through the ctypes wrappers of scoring.py / loss.py), with a fake (meta)
kernel for shape inference under ``torch.compile`` / ``torch.export`` and autograd registered
with ``register_autograd``.  There is no CPU kernel: the ops are registered for CUDA only.

    torch.ops.alad_b200.alignment_scores(im_set, s_seq, im_len, s_len, precision, aggregation) -> S
    torch.ops.alad_b200.alignment_scores_bwd(im_set, s_seq, im_len, s_len, aggregation, G) -> (d_im, d_s)
    torch.ops.alad_b200.dot_scores(im, s, precision) -> S
    torch.ops.alad_b200.triplet(scores, margin, max_violation) -> (loss, G)
    torch.ops.alad_b200.listnet(teacher, student) -> (loss, dM)
    torch.ops.alad_b200.rank_i2t(S, group, img_off) -> (rank, top1)
    torch.ops.alad_b200.rank_t2i(S, k, group) -> (rank, topk)

The nn.Module drop-ins of loss.py use torch.autograd.Function directly (they also accept CPU
inputs and move them); these ops are the traceable surface of the same kernels."""
from typing import List, Tuple

import torch
from torch import Tensor

from . import scoring

_lib = torch.library
NS = "alad_b200"


# ---------------------------------------------------------------------------------- alignment
@_lib.custom_op(f"{NS}::alignment_scores", mutates_args=(), device_types="cuda")
def alignment_scores(im_set: Tensor, s_seq: Tensor, im_len: List[int], s_len: List[int], precision: str,
                     aggregation: str) -> Tensor:
    return scoring.alignment_scores(im_set, s_seq, im_len, s_len, precision=precision, aggregation=aggregation)


@alignment_scores.register_fake
def _(im_set, s_seq, im_len, s_len, precision, aggregation):
    return im_set.new_empty((im_set.shape[0], s_seq.shape[0]), dtype=torch.float32)


@_lib.custom_op(f"{NS}::alignment_scores_bwd", mutates_args=(), device_types="cuda")
def alignment_scores_bwd(im_set: Tensor, s_seq: Tensor, im_len: List[int], s_len: List[int], aggregation: str,
                         G: Tensor) -> Tuple[Tensor, Tensor]:
    from . import loss as L
    im_c = scoring._require_cuda(im_set, "im_set")
    s_c = scoring._require_cuda(s_seq, "s_seq")
    _, W, nr, nw, _ = scoring.scored_counts(im_c.shape, s_c.shape, im_len, s_len)
    return L.alignment_backward(im_c, s_c, nr, nw, W, aggregation, G.float().contiguous())


@alignment_scores_bwd.register_fake
def _(im_set, s_seq, im_len, s_len, aggregation, G):
    return (im_set.new_empty(im_set.shape, dtype=torch.float32), s_seq.new_empty(s_seq.shape, dtype=torch.float32))


def _alignment_setup(ctx, inputs, output):
    im_set, s_seq, im_len, s_len, _, aggregation = inputs
    ctx.save_for_backward(im_set, s_seq)
    ctx.im_len, ctx.s_len, ctx.aggregation = list(im_len), list(s_len), aggregation


def _alignment_backward(ctx, g):
    im_set, s_seq = ctx.saved_tensors
    d_im, d_s = torch.ops.alad_b200.alignment_scores_bwd(im_set, s_seq, ctx.im_len, ctx.s_len, ctx.aggregation, g)
    return d_im, d_s, None, None, None, None


alignment_scores.register_autograd(_alignment_backward, setup_context=_alignment_setup)


# ---------------------------------------------------------------------------------- matching
@_lib.custom_op(f"{NS}::dot_scores", mutates_args=(), device_types="cuda")
def dot_scores(im: Tensor, s: Tensor, precision: str) -> Tensor:
    return scoring.dot_scores(im, s, precision=precision)


@dot_scores.register_fake
def _(im, s, precision):
    return im.new_empty((im.shape[0], s.shape[0]), dtype=torch.float32)


def _dot_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _dot_backward(ctx, g):
    im, s = ctx.saved_tensors
    g = g.contiguous().float()
    # G @ s and G.T @ im on the same tcgen05 GEMM (split precision)
    d_im = torch.ops.alad_b200.dot_scores(g, s.float().t().contiguous(), "fp32")
    d_s = torch.ops.alad_b200.dot_scores(g.t().contiguous(), im.float().t().contiguous(), "fp32")
    return d_im, d_s, None


dot_scores.register_autograd(_dot_backward, setup_context=_dot_setup)


# ---------------------------------------------------------------------------------- losses
@_lib.custom_op(f"{NS}::triplet", mutates_args=(), device_types="cuda")
def triplet(scores: Tensor, margin: float, max_violation: bool) -> Tuple[Tensor, Tensor]:
    from . import loss as L
    loss, G, _, _ = L.triplet_fwd_bwd(scores.float(), margin, max_violation, want_grad=True)
    return loss, G


@triplet.register_fake
def _(scores, margin, max_violation):
    return scores.new_empty((), dtype=torch.float32), scores.new_empty(scores.shape, dtype=torch.float32)


def _loss_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _triplet_backward(ctx, g_loss, g_G):
    (G,) = ctx.saved_tensors
    return G * g_loss, None, None


triplet.register_autograd(_triplet_backward, setup_context=_loss_setup)


@_lib.custom_op(f"{NS}::listnet", mutates_args=(), device_types="cuda")
def listnet(teacher: Tensor, student: Tensor) -> Tuple[Tensor, Tensor]:
    from . import loss as L
    loss, dM = L.listnet_fwd_bwd(teacher, student, want_grad=True)
    return loss, dM


@listnet.register_fake
def _(teacher, student):
    return student.new_empty((), dtype=torch.float32), student.new_empty(student.shape, dtype=torch.float32)


def _listnet_backward(ctx, g_loss, g_dM):
    (dM,) = ctx.saved_tensors
    return None, dM * g_loss                       # the teacher is detached (alad/loss.py:370)


listnet.register_autograd(_listnet_backward, setup_context=_loss_setup)


# ---------------------------------------------------------------------------------- ranking
@_lib.custom_op(f"{NS}::rank_i2t", mutates_args=(), device_types="cuda")
def rank_i2t(S: Tensor, group: int, img_off: int) -> Tuple[Tensor, Tensor]:
    from . import ranking
    return ranking.rank_rows(S, group, img_off)


@rank_i2t.register_fake
def _(S, group, img_off):
    return S.new_empty((S.shape[0],), dtype=torch.int32), S.new_empty((S.shape[0],), dtype=torch.int32)


@_lib.custom_op(f"{NS}::rank_t2i", mutates_args=(), device_types="cuda")
def rank_t2i(S: Tensor, k: int, group: int) -> Tuple[Tensor, Tensor]:
    from . import ranking
    return ranking.t2i_rank_topk(S, k, group)


@rank_t2i.register_fake
def _(S, k, group):
    return S.new_empty((S.shape[1],), dtype=torch.int32), S.new_empty((S.shape[1], k), dtype=torch.int32)


OPS = ("alignment_scores", "alignment_scores_bwd", "dot_scores", "triplet", "listnet", "rank_i2t", "rank_t2i")
