"""BASELINE config 5 (two-stage retrieval): matching-head top-K shortlist + alignment re-rank.
The reference has no such function; the oracle is the composition of its pieces (global-vector
mm of alad/recall_auxiliary.py:30 + MrSw scores of alad/loss.py:79-125 + descending sorts)."""
import numpy as np
import pytest
import torch

from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu


def _desc(v):
    return np.argsort(v, kind="stable")[::-1]


def _oracle_two_stage(images, captions, il, cl, K):
    ims = images[0::5]
    M = ims[:, 0, :].astype(np.float64) @ captions[:, 0, :].astype(np.float64).T
    S = O.mrsw_scores(ims, captions, il[0::5], cl, acc64=True).astype(np.float64)
    Ni, Nc = M.shape
    ri, rt = np.zeros(Ni), np.zeros(Nc)
    for i in range(Ni):
        o1 = _desc(M[i])
        short = o1[:K]
        order = np.concatenate([short[_desc(S[i, short])], o1[K:]])
        pos = np.empty(Nc, np.int64)
        pos[order] = np.arange(Nc)
        ri[i] = pos[5 * i:5 * i + 5].min()
    for c in range(Nc):
        o1 = _desc(M[:, c])
        short = o1[:K]
        order = np.concatenate([short[_desc(S[short, c])], o1[K:]])
        rt[c] = np.where(order == c // 5)[0][0]
    return ri, rt


@pytest.mark.parametrize("K", [10, 100])
def test_two_stage_matches_oracle_composition(K):
    from aladin_b200 import synth, two_stage
    images, captions, il, cl = synth.eval_containers(41, 120, 40, 128, max_regions=30, max_words=25, alpha=0.25)
    (m_i2t, m_t2i), (ri, rt) = two_stage.two_stage_retrieval(torch.from_numpy(images), torch.from_numpy(captions), il, cl,
                                                             shortlist=K, precision="fp32", return_ranks=True)
    ei, et = _oracle_two_stage(images, captions, il, cl, K)
    np.testing.assert_array_equal(ri, ei)
    np.testing.assert_array_equal(rt, et)
    np.testing.assert_allclose(m_i2t, O.recall_metrics(ei))
    np.testing.assert_allclose(m_t2i, O.recall_metrics(et))
