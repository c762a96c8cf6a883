"""CPU oracle for ALADIN's all-pairs cross-modal scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``aladin_b200/`` may import this module.
It is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the *checker*
and the *CPU baseline*, never as the thing shipped.

It is a numpy restatement (fp32 storage, BLAS matmul, optional fp64 accumulate)
of the reference algorithm; each function cites the reference lines it follows
(paths relative to the reference checkout, mesnico/ALADIN @ d1bd7bf).

Parity pinning: the reference ships **no** tests or golden vectors for this
path.  The oracle is therefore pinned against outputs of
the reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference (``alad.loss``, ``alad.evaluation``, ``alad.recall_auxiliary``) in the
authoring container and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """``F.normalize(x, p=2, dim=-1)`` (alad/loss.py:80-81): x / max(||x||, eps)."""
    x = np.asarray(x, dtype=F32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(-1, keepdims=True)).astype(F32)
    return (x / np.maximum(n, F32(eps))).astype(F32)


def l2norm_noeps(x: np.ndarray) -> np.ndarray:
    """``alad.utils.l2norm`` (alad/utils.py:134-139): no eps, rows, dim=1."""
    x = np.asarray(x, dtype=F32)
    n = np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x / n


def valid_count(length: int, extent: int) -> int:
    """Number of unmasked leading slots produced by ``mask[l:] = True`` on a row
    of ``extent`` slots (alad/loss.py:105-106, 111-112) -- Python slice semantics,
    including negative ``l`` (counts from the end) and ``l > extent``."""
    l = int(length)
    if l >= 0:
        return min(l, extent)
    return max(extent + l, 0)


def scored_extents(im_set_shape, s_seq_shape, im_len, s_len):
    """Slicing + length arithmetic of alad/loss.py:87-90.

    Returns (R, W, nr[Bi], nw[Bc]): container extents after dropping slot 0 of
    images and slot 0 / last two slots of captions, and the valid counts."""
    R = max(im_set_shape[1] - 1, 0)
    W = max(s_seq_shape[1] - 3, 0)
    nr = np.array([valid_count(l - 1, R) for l in im_len], dtype=np.int64)
    nw = np.array([valid_count(l - 3, W) for l in s_len], dtype=np.int64)
    return R, W, nr, nw


# --------------------------------------------------------------------------------------
# a1: alignment-head scores  (alad/loss.py:79-135)
# --------------------------------------------------------------------------------------

AGGREGATIONS = ("sum", "mean", "MrSw", "MrAVGw", "symm", "MwSr")


def alignment_tensor(im_set, s_seq, im_len, s_len):
    """Masked region x word cosine tensor A[Bi,Bc,R,W] (alad/loss.py:80-116).
    Only for small cases: materialises the 4-D tensor like the reference."""
    im = l2_normalize(im_set)[:, 1:, :]
    s = l2_normalize(s_seq)[:, 1:-2, :]
    R, W, nr, nw = scored_extents(im_set.shape, s_seq.shape, im_len, s_len)
    A = np.einsum("ird,jwd->ijrw", im, s, optimize=True).astype(F32)
    rmask = np.arange(R)[None, :] >= nr[:, None]          # [Bi,R] True = masked
    wmask = np.arange(W)[None, :] >= nw[:, None]          # [Bc,W]
    mask = rmask[:, None, :, None] | wmask[None, :, None, :]
    A[np.broadcast_to(mask, A.shape)] = 0.0
    return A, nr, nw


def pool_alignments(A, nw, aggregation="MrSw"):
    """Pooling modes of alad/loss.py:120-135 on a masked tensor A[Bi,Bc,R,W]."""
    if aggregation == "sum":
        return A.sum(axis=(2, 3), dtype=F32)
    if aggregation == "mean":
        return A.mean(axis=(2, 3), dtype=F32)
    if aggregation == "MrSw":
        return A.max(axis=2).sum(axis=2, dtype=F32)
    if aggregation == "MrAVGw":
        out = A.max(axis=2).sum(axis=2, dtype=F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (out / nw.astype(F32)[None, :]).astype(F32)
    if aggregation == "symm":
        return (A.max(axis=2).sum(axis=2, dtype=F32) + A.max(axis=3).sum(axis=2, dtype=F32)).astype(F32)
    if aggregation == "MwSr":
        return A.max(axis=3).sum(axis=2, dtype=F32)
    raise ValueError(f"unsupported aggregation {aggregation!r}")


def alignment_scores_small(im_set, s_seq, im_len, s_len, aggregation="MrSw"):
    """Reference-shaped evaluation (4-D tensor materialised); small inputs only."""
    A, nr, nw = alignment_tensor(im_set, s_seq, im_len, s_len)
    return pool_alignments(A, nw, aggregation)


def mrsw_scores(im_set, s_seq, im_len, s_len, chunk: int = 64, acc64: bool = False):
    """MrSw scores S[Bi,Bc] without the 4-D tensor (alad/loss.py:79-125).

    S[i,j] = sum_{w<nw_j} m(i,j,w),  m = max over the *padded* region extent, i.e.
    max(0, max_{r<nr_i} <im_i,r , s_j,w>) when nr_i < R and the plain max otherwise.
    One BLAS GEMM per image chunk: [chunk*R, d] x [d, Bc*W]."""
    im = l2_normalize(im_set)[:, 1:, :]
    s = l2_normalize(s_seq)[:, 1:-2, :]
    R, W, nr, nw = scored_extents(im_set.shape, s_seq.shape, im_len, s_len)
    Bi, Bc, d = im.shape[0], s.shape[0], im.shape[2]
    S = np.zeros((Bi, Bc), dtype=F32)
    if R == 0 or W == 0 or Bi == 0 or Bc == 0:
        return S
    wvalid = (np.arange(W)[None, :] < nw[:, None])                   # [Bc,W]
    s_flat = s.reshape(Bc * W, d)
    if acc64:
        s_flat = s_flat.astype(np.float64)
    for i0 in range(0, Bi, chunk):
        i1 = min(i0 + chunk, Bi)
        imc = im[i0:i1]
        if acc64:
            imc = imc.astype(np.float64)
        A = (imc.reshape(-1, d) @ s_flat.T).reshape(i1 - i0, R, Bc, W)   # [c,R,Bc,W]
        rmask = np.arange(R)[None, :] >= nr[i0:i1, None]                # [c,R]
        A = np.where(rmask[:, :, None, None], 0.0, A)
        m = A.max(axis=1)                                               # [c,Bc,W]
        m = np.where(wvalid[None], m, 0.0)
        S[i0:i1] = m.sum(axis=2).astype(F32)
    return S


def mrsw_scores_scalar(im_set, s_seq, im_len, s_len):
    """Pure-Python scalar restatement of SURVEY §3.3 steps 1-6 (tiny cases only)."""
    im_set = np.asarray(im_set, dtype=np.float64)
    s_seq = np.asarray(s_seq, dtype=np.float64)
    R, W, nr, nw = scored_extents(im_set.shape, s_seq.shape, im_len, s_len)
    Bi, Bc = im_set.shape[0], s_seq.shape[0]

    def unit(v):
        n = np.sqrt((v * v).sum())
        return v / max(n, 1e-12)

    S = np.zeros((Bi, Bc))
    for i in range(Bi):
        regs = [unit(im_set[i, 1 + r]) for r in range(R)]
        for j in range(Bc):
            tot = 0.0
            for w in range(W):
                if w >= nw[j]:
                    continue                       # masked word: whole column of zeros -> max 0
                sw = unit(s_seq[j, 1 + w])
                best = None
                for r in range(R):
                    a = float(regs[r] @ sw) if r < nr[i] else 0.0
                    best = a if best is None else max(best, a)
                tot += best if best is not None else 0.0
            S[i, j] = tot
    return S.astype(F32)


# --------------------------------------------------------------------------------------
# a3: matching-head scores (alad/loss.py:8-18)
# --------------------------------------------------------------------------------------


def dot_scores(im, s):
    """``dot_sim`` (alad/loss.py:8-11)."""
    return (np.asarray(im, F32) @ np.asarray(s, F32).T).astype(F32)


def cosine_scores(im, s):
    """``cosine_sim`` (alad/loss.py:13-18): l2norm without eps, then mm."""
    return dot_scores(l2norm_noeps(im), l2norm_noeps(s))


# --------------------------------------------------------------------------------------
# a2: hinge triplet loss (alad/loss.py:42-67) and its gradient
# --------------------------------------------------------------------------------------


def triplet_loss(S, margin: float, max_violation: bool):
    S = np.asarray(S, dtype=F32)
    B = S.shape[0]
    assert S.shape == (B, B), "compute_contrastive_loss needs a square matrix (torch.eye, loss.py:55)"
    diag = np.diag(S).astype(F32)
    cost_s = np.maximum(F32(margin) + S - diag[:, None], F32(0))      # rows: caption retrieval
    cost_im = np.maximum(F32(margin) + S - diag[None, :], F32(0))     # cols: image retrieval
    eye = np.eye(B, dtype=bool)
    cost_s[eye] = 0
    cost_im[eye] = 0
    if max_violation:
        return F32(cost_s.max(axis=1).sum(dtype=F32) + cost_im.max(axis=0).sum(dtype=F32))
    return F32(cost_s.sum(dtype=F32) + cost_im.sum(dtype=F32))


def triplet_grad(S, margin: float, max_violation: bool):
    """dL/dS of :func:`triplet_loss` (first-occurrence arg-max, like torch CPU)."""
    S = np.asarray(S, dtype=F32)
    B = S.shape[0]
    diag = np.diag(S).astype(F32)
    cost_s = np.maximum(F32(margin) + S - diag[:, None], F32(0))
    cost_im = np.maximum(F32(margin) + S - diag[None, :], F32(0))
    eye = np.eye(B, dtype=bool)
    cost_s[eye] = 0
    cost_im[eye] = 0
    G = np.zeros((B, B), dtype=F32)
    if max_violation:
        for i in range(B):
            j = int(cost_s[i].argmax())
            if cost_s[i, j] > 0:
                G[i, j] += 1
                G[i, i] -= 1
        for j in range(B):
            i = int(cost_im[:, j].argmax())
            if cost_im[i, j] > 0:
                G[i, j] += 1
                G[j, j] -= 1
    else:
        ps = (cost_s > 0).astype(F32)
        pi = (cost_im > 0).astype(F32)
        G = ps + pi
        G[eye] = -(ps.sum(axis=1) + pi.sum(axis=0))
    return G


# --------------------------------------------------------------------------------------
# a4: ListNet distillation (alad/loss.py:369-370, 427-445) and gradient
# --------------------------------------------------------------------------------------


def _softmax(x, axis):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def listnet_loss(teacher, student, temperature: float = 6.0, eps: float = 1e-10):
    T = np.asarray(teacher, dtype=F32)
    M = np.asarray(student, dtype=F32)
    total = F32(0)
    for axis in (0, 1):                       # im_cost (dim=0) + s_cost (dim=1)
        p = _softmax(M * F32(temperature), axis).astype(F32)
        t = _softmax(T, axis).astype(F32)
        cost = -(t * np.log(p + F32(eps))).sum(axis=axis, dtype=F32)
        total = F32(total + cost.mean(dtype=F32))
    return total


def listnet_grad(teacher, student, temperature: float = 6.0, eps: float = 1e-10):
    """dL/dstudent in fp64 (teacher is detached: alad/loss.py:370)."""
    T = np.asarray(teacher, dtype=np.float64)
    M = np.asarray(student, dtype=np.float64)
    G = np.zeros_like(M)
    for axis in (0, 1):
        n = M.shape[axis]
        p = _softmax(M * temperature, axis)
        t = _softmax(T, axis)
        a = t * p / (p + eps)
        G += (temperature / n) * (p * a.sum(axis=axis, keepdims=True) - a)
    return G.astype(F32)


# --------------------------------------------------------------------------------------
# DistillationLoss, remaining modes (alad/loss.py:371-425) and their gradients
# --------------------------------------------------------------------------------------


def distill_mse(teacher, student, wb):
    """mode 'mse' (alad/loss.py:371-373): mean((student*wb0 + wb1 - teacher)^2) over all B*B entries.
    Returns (loss, d_student, d_wb) -- gradients in fp64, cast to fp32."""
    T = np.asarray(teacher, dtype=np.float64)
    M = np.asarray(student, dtype=np.float64)
    w0, w1 = float(wb[0]), float(wb[1])
    diff = M * w0 + w1 - T
    n = diff.size
    loss = (diff * diff).sum() / n
    dM = (2.0 * w0 / n) * diff
    dwb = np.array([(2.0 / n) * (diff * M).sum(), (2.0 / n) * diff.sum()])
    return F32(loss), dM.astype(F32), dwb.astype(F32)


def distill_contrastive(teacher, student, margin: float = 0.2):
    """mode 'contrastive' (alad/loss.py:397-418).  The hard negatives come from the teacher with
    its diagonal zeroed: ns[k] = argmax_j Tnd[k, j], ni[k] = argmax_i Tnd[i, k].  The reference then
    index_selects WHOLE columns / rows of the (un-cleared) hinge matrices and sums everything:
      loss = sum_i sum_k [m + M[i, ns[k]] - M[i, i]]_+  +  sum_k sum_j [m + M[ni[k], j] - M[j, j]]_+
    (so the diagonal terms contribute the margin itself).  Returns (loss, d_student)."""
    T = np.array(teacher, dtype=F32, copy=True)
    M = np.asarray(student, dtype=F32)
    B = T.shape[0]
    T[np.eye(B, dtype=bool)] = 0
    ns = T.argmax(axis=1)                      # first occurrence, like torch CPU
    ni = T.argmax(axis=0)
    diag = np.diag(M).astype(F32)
    cost_s = np.maximum(F32(margin) + M - diag[:, None], F32(0))
    cost_im = np.maximum(F32(margin) + M - diag[None, :], F32(0))
    sel_s = cost_s[:, ns]                      # [B, B]: column ns[k] for every k
    sel_im = cost_im[ni, :]
    loss = F32(sel_s.sum(dtype=F32) + sel_im.sum(dtype=F32))
    G = np.zeros((B, B), dtype=np.float64)
    for k in range(B):
        a = cost_s[:, ns[k]] > 0               # rows i with an active hinge on column ns[k]
        G[a, ns[k]] += 1
        G[np.arange(B)[a], np.arange(B)[a]] -= 1
        b = cost_im[ni[k], :] > 0              # columns j with an active hinge on row ni[k]
        G[ni[k], b] += 1
        G[np.arange(B)[b], np.arange(B)[b]] -= 1
    return loss, G.astype(F32)


def distill_ordinal(teacher, student, margin: float = 0.2, threshold: float = 0.1, stride: int = 3):
    """mode 'ordinal' (alad/loss.py:374-396): every row (column) of the student is reordered by the
    ascending teacher order; entries `stride` apart must keep that order by `margin`, counted only
    where the LATER teacher value is >= threshold.  Each direction is a mean over its selected
    differences (NaN when nothing is selected, like torch's mean of an empty tensor).
    Returns (loss, d_student)."""
    T = np.asarray(teacher, dtype=F32)
    M = np.asarray(student, dtype=np.float64)
    G = np.zeros_like(M)
    total = 0.0
    for axis in (1, 0):
        Tt = T if axis == 1 else T.T
        Mt = M if axis == 1 else M.T
        Gt = np.zeros_like(Mt)
        order = np.argsort(Tt, axis=1, kind="stable")
        Ts = np.take_along_axis(Tt, order, axis=1)
        Ms = np.take_along_axis(Mt, order, axis=1)
        diff = Ms[:, :-stride] - Ms[:, stride:]
        valid = Ts[:, stride:] >= F32(threshold)
        n = int(valid.sum())
        if n == 0:
            total = float("nan")
            continue
        h = margin + diff
        total += np.where(valid, np.maximum(h, 0.0), 0.0).sum() / n
        act = valid & (h > 0)
        rows, pos = np.nonzero(act)
        np.add.at(Gt, (rows, order[rows, pos]), 1.0 / n)
        np.add.at(Gt, (rows, order[rows, pos + stride]), -1.0 / n)
        G += Gt if axis == 1 else Gt.T
    return F32(total), G.astype(F32)


def order_scores(im, s):
    """``order_sim`` (alad/loss.py:20-26): score[i, j] = -|| max(0, s_j - im_i) ||_2."""
    im = np.asarray(im, F32)
    s = np.asarray(s, F32)
    y = np.maximum(s[None, :, :] - im[:, None, :], F32(0))
    return (-np.sqrt((y * y).sum(axis=2, dtype=F32))).astype(F32)


def cosine_backward(im, s, G):
    """Gradient of sum(G * cosine_sim(im, s)) w.r.t. the raw inputs (alad/loss.py:13-18 through
    alad.utils.l2norm, no eps), fp64 internally."""
    im = np.asarray(im, np.float64)
    s = np.asarray(s, np.float64)
    G = np.asarray(G, np.float64)
    ni = np.sqrt((im * im).sum(1, keepdims=True))
    ns = np.sqrt((s * s).sum(1, keepdims=True))
    ih, sh = im / ni, s / ns
    d_ih, d_sh = G @ sh, G.T @ ih
    d_im = (d_ih - ih * (ih * d_ih).sum(1, keepdims=True)) / ni
    d_s = (d_sh - sh * (sh * d_sh).sum(1, keepdims=True)) / ns
    return d_im.astype(F32), d_s.astype(F32)


def pooled_sum_backward(im_set, s_seq, im_len, s_len, G, mean: bool = False):
    """Gradient of sum(G * S) for aggregation 'sum' / 'mean' (alad/loss.py:120-123):
    S[i,j] = <sum_r imhat(i,r), sum_w shat(j,w)> over the valid tokens (/ (R*W) for 'mean')."""
    im_raw = np.asarray(im_set, dtype=np.float64)
    s_raw = np.asarray(s_seq, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    R, W, nr, nw = scored_extents(im_raw.shape, s_raw.shape, im_len, s_len)
    if mean:
        G = G / max(R * W, 1)

    def norm(x):
        n = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)
        return x / n, n

    imh, imn = norm(im_raw)
    sh, sn = norm(s_raw)
    rv = (np.arange(im_raw.shape[1])[None, :] >= 1) & (np.arange(im_raw.shape[1])[None, :] < 1 + nr[:, None])
    wv = (np.arange(s_raw.shape[1])[None, :] >= 1) & (np.arange(s_raw.shape[1])[None, :] < 1 + nw[:, None])
    pi = (imh * rv[:, :, None]).sum(1)          # [Bi,d]
    ps = (sh * wv[:, :, None]).sum(1)           # [Bc,d]
    d_pi, d_ps = G @ ps, G.T @ pi
    d_imh = rv[:, :, None] * d_pi[:, None, :]
    d_sh = wv[:, :, None] * d_ps[:, None, :]

    def norm_bwd(xh, n, dxh):
        return (dxh - xh * (xh * dxh).sum(-1, keepdims=True)) / n

    return norm_bwd(imh, imn, d_imh).astype(F32), norm_bwd(sh, sn, d_sh).astype(F32)


# --------------------------------------------------------------------------------------
# MrSw backward: d im_set, d s_seq from G = dL/dS
# --------------------------------------------------------------------------------------


def mrsw_backward(im_set, s_seq, im_len, s_len, G):
    """Gradient of sum(G * S) w.r.t. the raw (un-normalised) inputs, fp64 internally.
    Follows autograd through masked_fill_ / max / F.normalize (alad/loss.py:80-125)."""
    im_raw = np.asarray(im_set, dtype=np.float64)
    s_raw = np.asarray(s_seq, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    R, W, nr, nw = scored_extents(im_raw.shape, s_raw.shape, im_len, s_len)

    def norm(x):
        n = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)
        return x / n, n

    imh, imn = norm(im_raw)
    sh, sn = norm(s_raw)
    d_imh = np.zeros_like(imh)
    d_sh = np.zeros_like(sh)
    Bi, Bc = imh.shape[0], sh.shape[0]
    for i in range(Bi):
        if R == 0:
            break
        regs = imh[i, 1:1 + R]                                   # [R,d]
        for j in range(Bc):
            g = G[i, j]
            if g == 0.0 or nw[j] == 0:
                continue
            words = sh[j, 1:1 + nw[j]]                           # [nw,d]
            A = regs @ words.T                                   # [R,nw]
            A[nr[i]:, :] = 0.0
            rstar = A.argmax(axis=0)                             # first occurrence, like torch CPU max
            for w in range(nw[j]):
                r = int(rstar[w])
                if r < nr[i]:                                    # winner unmasked -> gradient flows
                    d_imh[i, 1 + r] += g * words[w]
                    d_sh[j, 1 + w] += g * regs[r]

    def norm_bwd(xh, n, dxh):
        return (dxh - xh * (xh * dxh).sum(-1, keepdims=True)) / n

    return norm_bwd(imh, imn, d_imh).astype(F32), norm_bwd(sh, sn, d_sh).astype(F32)


# --------------------------------------------------------------------------------------
# aggregation 'scan-sentences' (alad/loss.py:136-149): attention of every region over the words
# --------------------------------------------------------------------------------------


def _scan_pair_forward(C, K):
    """One (image, caption) pair on its VALID tokens.  C[nr,nw] = region x word cosines, K[nw,nw] = Gram matrix of
    the caption's unit word vectors.  Follows alad/loss.py:137-147: relu, L2-normalise over the regions (dim=2,
    eps 1e-12), softmax over the unmasked words (masked logits are -inf), attended word vector per region, cosine
    of region and attended vector (eps 1e-8; <x, att> = alpha.C and ||att||^2 = alpha' K alpha).  Returns the
    intermediates needed by the backward."""
    P = np.maximum(C, 0.0)
    n = np.sqrt((P * P).sum(axis=0))                                   # [nw] norm over the regions
    inv_n = 1.0 / np.maximum(n, 1e-12)
    Q = P * inv_n[None, :]
    E = np.exp(Q - Q.max(axis=1, keepdims=True))
    alpha = E / E.sum(axis=1, keepdims=True)                           # [nr,nw]
    T = alpha @ K                                                      # [nr,nw]  (K symmetric)
    u = (alpha * C).sum(axis=1)
    v = (alpha * T).sum(axis=1)
    b = np.sqrt(np.maximum(v, 0.0))
    new = u / np.maximum(b, 1e-8)
    return dict(P=P, n=n, inv_n=inv_n, Q=Q, alpha=alpha, T=T, u=u, v=v, b=b, new=new)


def _scan_pair_backward(C, K, g):
    """(dL/dC [nr,nw], dL/dK [nw,nw]) of one pair for dL/dS = g (autograd through alad/loss.py:137-149)."""
    f = _scan_pair_forward(C, K)
    alpha, T, u, b = f["alpha"], f["T"], f["u"], f["b"]
    live = b > 1e-8                                                   # clamp of F.cosine_similarity inactive
    cu = np.where(live, g / np.maximum(b, 1e-300), g / 1e-8)          # g * d new / d u
    cv = np.where(live, -0.5 * g * u / np.maximum(b, 1e-300) ** 3, 0.0)   # g * d new / d v
    d_alpha = cu[:, None] * C + 2.0 * cv[:, None] * T
    dQ = alpha * (d_alpha - (alpha * d_alpha).sum(axis=1, keepdims=True))
    dot = (f["Q"] * dQ).sum(axis=0)
    dP = np.where(f["n"][None, :] > 1e-12, (dQ - f["Q"] * dot[None, :]) * f["inv_n"][None, :], dQ * 1e12)
    dC = dP * (C > 0.0) + cu[:, None] * alpha
    return dC, (alpha * cv[:, None]).T @ alpha


def scan_scores(im_set, s_seq, im_len, s_len, acc64: bool = True):
    """Scores of aggregation 'scan-sentences' (alad/loss.py:136-149) without the 4-D tensors.
    Masked regions contribute 0 (loss.py:147); a caption without valid words gives NaN for every image that
    has valid regions (softmax over an all -inf row, loss.py:139-140), like the reference."""
    dt = np.float64 if acc64 else F32
    imh = l2_normalize(im_set).astype(dt)
    sh = l2_normalize(s_seq).astype(dt)
    R, W, nr, nw = scored_extents(im_set.shape, s_seq.shape, im_len, s_len)
    if W == 0:
        raise IndexError("scan-sentences indexes word 0 of the mask (alad/loss.py:147): empty word extent")
    Bi, Bc = imh.shape[0], sh.shape[0]
    S = np.zeros((Bi, Bc), dtype=dt)
    for j in range(Bc):
        Y = sh[j, 1:1 + nw[j]]
        K = Y @ Y.T
        for i in range(Bi):
            if nr[i] == 0:
                continue
            if nw[j] == 0:
                S[i, j] = np.nan
                continue
            X = imh[i, 1:1 + nr[i]]
            S[i, j] = _scan_pair_forward(X @ Y.T, K)["new"].sum()
    return S.astype(F32)


def scan_fragile_pairs(im_set, s_seq, im_len, s_len, tol=3e-5, floor=1e-3):
    """bool [Bi,Bc]: pairs on which 'scan-sentences' is DISCONTINUOUS in the cosines.  alad/loss.py:137-138 feeds
    relu(cos) through F.normalize over the regions: when every other region of a word column is <= 0 (column norm
    without it < floor) and one cosine sits within `tol` of 0, that entry normalises to 0 or to 1 depending on its
    sign, so two fp32-grade GEMMs (the reference's own CPU and GPU matmuls included) legitimately disagree there.
    Parity tests assert the tight bound on all other pairs and a loose one (|dS| <= 1, one region's cosine) here."""
    imh = l2_normalize(im_set).astype(np.float64)
    sh = l2_normalize(s_seq).astype(np.float64)
    _, _, nr, nw = scored_extents(im_set.shape, s_seq.shape, im_len, s_len)
    out = np.zeros((imh.shape[0], sh.shape[0]), dtype=bool)
    for j in range(sh.shape[0]):
        Y = sh[j, 1:1 + nw[j]]
        for i in range(imh.shape[0]):
            if nr[i] == 0 or nw[j] == 0:
                continue
            C = imh[i, 1:1 + nr[i]] @ Y.T
            near = np.abs(C) < tol
            rest = np.sqrt((np.where(near, 0.0, np.maximum(C, 0.0)) ** 2).sum(axis=0))
            out[i, j] = bool((near.any(axis=0) & (rest < floor)).any())
    return out


def scan_backward(im_set, s_seq, im_len, s_len, G):
    """Gradient of sum(G * S) for aggregation 'scan-sentences' w.r.t. the raw inputs, fp64 internally
    (autograd through alad/loss.py:80-81, 137-149).  Pairs with nw = 0 are skipped (the reference yields NaN)."""
    im_raw = np.asarray(im_set, dtype=np.float64)
    s_raw = np.asarray(s_seq, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    R, W, nr, nw = scored_extents(im_raw.shape, s_raw.shape, im_len, s_len)

    def norm(x):
        n = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)
        return x / n, n

    imh, imn = norm(im_raw)
    sh, sn = norm(s_raw)
    d_imh = np.zeros_like(imh)
    d_sh = np.zeros_like(sh)
    for j in range(sh.shape[0]):
        if nw[j] == 0:
            continue
        Y = sh[j, 1:1 + nw[j]]
        K = Y @ Y.T
        dK = np.zeros_like(K)
        for i in range(imh.shape[0]):
            g = G[i, j]
            if g == 0.0 or nr[i] == 0:
                continue
            X = imh[i, 1:1 + nr[i]]
            dC, dKp = _scan_pair_backward(X @ Y.T, K, g)
            dK += dKp
            d_imh[i, 1:1 + nr[i]] += dC @ Y
            d_sh[j, 1:1 + nw[j]] += dC.T @ X
        d_sh[j, 1:1 + nw[j]] += (dK + dK.T) @ Y

    def norm_bwd(xh, n, dxh):
        return (dxh - xh * (xh * dxh).sum(-1, keepdims=True)) / n

    return norm_bwd(imh, imn, d_imh).astype(F32), norm_bwd(sh, sn, d_sh).astype(F32)


# --------------------------------------------------------------------------------------
# a6-a8: ranking (alad/evaluation.py:213-235, 303-320; alad/recall_auxiliary.py:34-64)
# --------------------------------------------------------------------------------------


def i2t_ranks(S):
    """S[Ni,Nc] image-major scores, captions 5i..5i+4 are image i's ground truth.
    rank_i = best position of a GT caption in the descending order (evaluation.py:213-223)."""
    S = np.asarray(S)
    Ni = S.shape[0]
    ranks = np.zeros(Ni)
    top1 = np.zeros(Ni)
    for i in range(Ni):
        inds = np.argsort(S[i])[::-1]
        pos = np.empty_like(inds)
        pos[inds] = np.arange(inds.size)
        ranks[i] = pos[5 * i:5 * i + 5].min()
        top1[i] = inds[0]
    return ranks, top1


def t2i_ranks(S, k: int = 50):
    """rank_c = position of image c//5 in caption c's descending order; top-k image
    indices per caption (evaluation.py:303-308; recall_auxiliary.py:52-56 with k=1)."""
    S = np.asarray(S)
    Ni, Nc = S.shape
    ranks = np.zeros(Nc)
    topk = np.zeros((Nc, k))
    for c in range(Nc):
        inds = np.argsort(S[:, c])[::-1]
        ranks[c] = np.where(inds == c // 5)[0][0]
        topk[c] = inds[:k]
    return ranks, topk


def recall_metrics(ranks):
    """R@1/5/10, medr, meanr (evaluation.py:231-235)."""
    ranks = np.asarray(ranks)
    n = len(ranks)
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / n
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / n
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / n
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return r1, r5, r10, medr, meanr


def i2t(images, captions, img_lens, cap_lens, npts=None, cap_batches=1, use_alignment=True):
    """Per-query restatement of ``i2t`` (alad/evaluation.py:158-241): one query image
    (row 5i) against the gallery in ``cap_batches`` chunks, argsort, rank of the GT."""
    images = np.asarray(images, F32)
    captions = np.asarray(captions, F32)
    if npts is None:
        npts = images.shape[0] // 5
    per = captions.shape[0] // cap_batches
    ranks = np.zeros(npts)
    top1 = np.zeros(npts)
    for index in range(npts):
        im = images[5 * index][None]
        if use_alignment:
            parts = [mrsw_scores(im, captions[b * per:(b + 1) * per], [img_lens[5 * index]],
                                 cap_lens[b * per:(b + 1) * per]).ravel() for b in range(cap_batches)]
            d = np.concatenate(parts)
        else:
            d = (im[:, 0, :] @ captions[:, 0, :].T).ravel()
        inds = np.argsort(d)[::-1]
        pos = np.empty_like(inds)
        pos[inds] = np.arange(inds.size)
        ranks[index] = pos[5 * index:5 * index + 5].min()
        top1[index] = inds[0]
    return recall_metrics(ranks) + (0, 0), (ranks, top1)


def t2i(images, captions, img_lens, cap_lens, npts=None, im_batches=1, use_alignment=True):
    """Per-query-group restatement of ``t2i`` (alad/evaluation.py:244-327)."""
    images = np.asarray(images, F32)
    captions = np.asarray(captions, F32)
    if npts is None:
        npts = images.shape[0] // 5
    ims = images[0::5]
    ims_len = [img_lens[i] for i in range(0, images.shape[0], 5)]
    per = ims.shape[0] // im_batches
    ranks = np.zeros(5 * npts)
    top50 = np.zeros((5 * npts, 50))
    for index in range(npts):
        q = captions[5 * index:5 * index + 5]
        qlen = cap_lens[5 * index:5 * index + 5]
        if use_alignment:
            parts = [mrsw_scores(ims[b * per:(b + 1) * per], q, ims_len[b * per:(b + 1) * per], qlen).T
                     for b in range(im_batches)]
            d = np.concatenate(parts, axis=1)
        else:
            d = q[:, 0, :] @ ims[:, 0, :].T
        for i in range(d.shape[0]):
            inds = np.argsort(d[i])[::-1]
            ranks[5 * index + i] = np.where(inds == index)[0][0]
            top50[5 * index + i] = inds[:50]
    return recall_metrics(ranks) + (0, 0), (ranks, top50)


def compute_recall(img_embs, cap_embs):
    """``compute_recall`` / ``recall_test`` / ``recall`` (alad/recall_auxiliary.py:8-88,133-149)
    on global vectors: returns (r1,r5,r10,r1i,r5i,r10i,rsum)."""
    ims = np.asarray(img_embs, F32)[0::5]
    caps = np.asarray(cap_embs, F32)
    S = ims @ caps.T
    ri, _ = i2t_ranks(S)
    rt, _ = t2i_ranks(S, k=1)
    r1, r5, r10, _, _ = recall_metrics(ri)
    r1i, r5i, r10i, _, _ = recall_metrics(rt)
    return r1, r5, r10, r1i, r5i, r10i, r1 + r5 + r10 + r1i + r5i + r10i
