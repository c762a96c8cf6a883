"""CPU baseline port of the reference's evaluation path in torch CPU ops.

TEST / BENCH INFRASTRUCTURE ONLY (same rules as alad_oracle.py: never imported by aladin_b200/).

Why a second restatement: the reference IS PyTorch, and its CPU speed comes from torch's threaded
kernels (F.normalize, batched matmul on expanded operands, masked_fill_, max, sum).  The numpy oracle
is exact but 4-8x slower than the reference itself on the same cores for the i2t direction, which
would flatter every GPU/CPU ratio.  This module follows the reference op for op so that
`bench.py --impl reference` / `cpu_baseline` time what the reference would cost on the box:

  alignment_scores()  ~ AlignmentContrastiveLoss.forward, alad/loss.py:79-125 ('MrSw', no loss)
  i2t() / t2i()       ~ alad/evaluation.py:158-241 / 244-327 (per-query loops, numpy argsort)

Pinned by tests/test_oracle_golden.py against the same golden vectors as the numpy oracle, and
against the numpy oracle itself."""
import numpy as np
import torch
import torch.nn.functional as F


def _length_mask(lengths, extent):
    """True = masked; row b is masked from slot lengths[b] on (python slice semantics, loss.py:103-112)."""
    m = torch.zeros(len(lengths), extent, dtype=torch.bool)
    for row, l in zip(m, lengths):
        row[l:] = True
    return m


def alignment_scores(im_set, s_seq, im_len, s_len):
    """S[Bi,Bc] = sum over words of the max over regions of the masked cosines (alad/loss.py:79-125)."""
    im = F.normalize(im_set, p=2, dim=2)[:, 1:, :]
    s = F.normalize(s_seq, p=2, dim=2)[:, 1:-2, :]
    Bi, R = im.shape[0], im.shape[1]
    Bc, W = s.shape[0], s.shape[1]
    # the reference expands both operands to [Bi,Bc,*,d] and lets matmul batch over the pairs (loss.py:97-99)
    im_e = im.unsqueeze(1).expand(-1, Bc, -1, -1)
    s_e = s.unsqueeze(0).expand(Bi, -1, -1, -1)
    A = torch.matmul(im_e, s_e.permute(0, 1, 3, 2))                     # [Bi,Bc,R,W]
    rmask = _length_mask([l - 1 for l in im_len], R)
    wmask = _length_mask([l - 3 for l in s_len], W)
    mask = rmask.unsqueeze(2).unsqueeze(1).expand(-1, Bc, -1, W) | wmask.unsqueeze(1).unsqueeze(0).expand(Bi, -1, R, -1)
    A.masked_fill_(mask, 0)
    return A.max(2)[0].sum(2)


def _metrics(ranks):
    n = len(ranks)
    return (100.0 * np.count_nonzero(ranks < 1) / n, 100.0 * np.count_nonzero(ranks < 5) / n,
            100.0 * np.count_nonzero(ranks < 10) / n, np.floor(np.median(ranks)) + 1, ranks.mean() + 1)


def i2t(images, captions, img_lens, cap_lens, npts=None, cap_batches=1):
    """One query image (row 5i) at a time against the caption gallery in `cap_batches` chunks."""
    images, captions = torch.as_tensor(images), torch.as_tensor(captions)
    if npts is None:
        npts = images.shape[0] // 5
    per = captions.shape[0] // cap_batches
    ranks, top1 = np.zeros(npts), np.zeros(npts)
    with torch.no_grad():
        for q in range(npts):
            im = images[5 * q:5 * q + 1]
            parts = [alignment_scores(im, captions[b * per:(b + 1) * per], [img_lens[5 * q]], cap_lens[b * per:(b + 1) * per])
                     for b in range(cap_batches)]
            d = torch.cat(parts, dim=1).numpy().ravel()
            order = np.argsort(d)[::-1]
            pos = np.empty_like(order)
            pos[order] = np.arange(order.size)
            ranks[q] = pos[5 * q:5 * q + 5].min()
            top1[q] = order[0]
    return _metrics(ranks) + (0, 0), (ranks, top1)


def t2i(images, captions, img_lens, cap_lens, npts=None, im_batches=1):
    """Five query captions at a time against the distinct gallery images (rows 0::5)."""
    images, captions = torch.as_tensor(images), torch.as_tensor(captions)
    if npts is None:
        npts = images.shape[0] // 5
    ims = images[0::5]
    ims_len = [img_lens[i] for i in range(0, images.shape[0], 5)]
    per = ims.shape[0] // im_batches
    ranks, top50 = np.zeros(5 * npts), np.zeros((5 * npts, 50))
    with torch.no_grad():
        for q in range(npts):
            caps = captions[5 * q:5 * q + 5]
            lens = cap_lens[5 * q:5 * q + 5]
            parts = [alignment_scores(ims[b * per:(b + 1) * per], caps, ims_len[b * per:(b + 1) * per], lens).t()
                     for b in range(im_batches)]
            d = torch.cat(parts, dim=1).numpy()
            for j in range(d.shape[0]):
                order = np.argsort(d[j])[::-1]
                ranks[5 * q + j] = np.where(order == q)[0][0]
                top50[5 * q + j] = order[:50]
    return _metrics(ranks) + (0, 0), (ranks, top50)
