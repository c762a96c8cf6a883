"""Drop-ins for ``alad.recall_auxiliary`` (alad/recall_auxiliary.py:8-149): recall of the
matching head from global vectors.  One tcgen05 GEMM (ims @ caps.T) + the ranking kernels
instead of ``mm`` + per-query numpy argsort loops."""
import numpy as np
import torch

from . import ranking, retrieval, scoring


def recall(images, captions, model=None, mode='i2t', lenghts=None, return_ranks=False):
    """images/captions [N,d]; image rows 0::5 are the distinct images (recall_auxiliary.py:14-27)."""
    ims = images[0::5]
    S = scoring.dot_scores(ims, captions)
    if mode == 'i2t':
        rank, top1 = ranking.rank_rows(S, 5, 0)
    elif mode == 't2i':
        rank, top = ranking.t2i_rank_topk(S, 1)
        top1 = top[:, 0]
    else:
        raise ValueError('mode not correct')
    ranks = rank.cpu().numpy().astype(np.float64)
    top1 = top1.cpu().numpy().astype(np.float64)
    r1, r5, r10, medr, meanr = retrieval.recall_tuple(ranks)
    if return_ranks:
        return (r1, r5, r10, medr, meanr), (ranks, top1)
    return r1, r5, r10, medr, meanr


def recall_test(img_embs, cap_embs, tot_lengths=None, model=None):
    with torch.no_grad():
        ims = img_embs[0::5]
        S = scoring.dot_scores(ims, cap_embs)                       # one GEMM serves both directions
        ri, _, rt, _, _ = ranking.rank_fused(S, 1)                # both directions: two sweeps of S
        r1, r5, r10, _, _ = retrieval.recall_tuple(ri.cpu().numpy().astype(np.float64))
        r1i, r5i, r10i, _, _ = retrieval.recall_tuple(rt.cpu().numpy().astype(np.float64))
    return r1, r5, r10, r1i, r5i, r10i, r1 + r5 + r10 + r1i + r5i + r10i


FOLD_ROWS = 5000          # rows per fold: 1000 images x 5 (alad/recall_auxiliary.py:99-100)


def recall_1k_5fold_test(img_embs, cap_embs, tot_lengths=None, model=None):
    """alad/recall_auxiliary.py:90-130: mean of recall_test over the first five 5000-row chunks (IndexError, like
    the reference, when there are fewer than five)."""
    img_chunks = torch.split(img_embs, FOLD_ROWS, dim=0)
    cap_chunks = torch.split(cap_embs, FOLD_ROWS, dim=0)
    folds = []
    for i in range(5):
        print('Computing Test recall... chunk %s of 5' % (i + 1))
        folds.append(recall_test(img_chunks[i], cap_chunks[i]))
    r1, r5, r10, r1i, r5i, r10i = (float(np.mean([f[j] for f in folds])) for j in range(6))
    rsum = r1 + r5 + r10 + r1i + r5i + r10i
    print("Test 1K 5Folds - Recall Image to text: %.2f, %.2f, %.2f" % (r1, r5, r10))
    print("Test 1K 5Folds - Recall Text to image: %.2f, %.2f, %.2f" % (r1i, r5i, r10i))
    print('Test 1K 5Folds - Sum score: %.2f' % rsum)
    return r1, r5, r10, r1i, r5i, r10i, rsum


def compute_recall(img_embs, cap_embs, tot_lengths=None, model=None):
    r1, r5, r10, r1i, r5i, r10i, rsum = recall_test(img_embs, cap_embs)
    print("Recall Image to text: %.2f, %.2f, %.2f" % (r1, r5, r10))
    print("Recall Text to image: %.2f, %.2f, %.2f" % (r1i, r5i, r10i))
    print('Sum score: %.2f' % rsum)
    return r1, r5, r10, r1i, r5i, r10i, rsum
