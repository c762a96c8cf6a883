"""GPU parity tests for the ranking kernels against the oracle (bit-exact integers)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu


def _stable_desc_order(v):
    return np.argsort(v, kind="stable")[::-1]


@pytest.mark.parametrize("Ni,ties", [(60, False), (333, True), (1000, False)])
def test_rank_kernels_match_oracle(Ni, ties):
    from aladin_b200 import ranking
    r = np.random.RandomState(Ni)
    Nc = 5 * Ni
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    S[np.arange(Nc) // 5, np.arange(Nc)] += 1.5
    if ties:
        S = np.round(S * 2) / 2                      # many exact ties
    Sd = torch.from_numpy(S).cuda()
    rank, top1 = ranking.rank_rows(Sd)
    count, top50 = ranking.t2i_rank_topk(Sd, 50)
    torch.cuda.synchronize()
    # expected under the documented total order (score desc, index desc on ties)
    exp_rank = np.zeros(Ni, np.int64); exp_top1 = np.zeros(Ni, np.int64)
    for i in range(Ni):
        inds = _stable_desc_order(S[i])
        pos = np.empty(Nc, np.int64); pos[inds] = np.arange(Nc)
        exp_rank[i] = pos[5 * i:5 * i + 5].min(); exp_top1[i] = inds[0]
    exp_rt = np.zeros(Nc, np.int64); exp_top50 = np.zeros((Nc, 50), np.int64)
    for c in range(Nc):
        inds = _stable_desc_order(S[:, c])
        exp_rt[c] = np.where(inds == c // 5)[0][0]; exp_top50[c] = inds[:50]
    np.testing.assert_array_equal(rank.cpu().numpy(), exp_rank)
    np.testing.assert_array_equal(top1.cpu().numpy(), exp_top1)
    np.testing.assert_array_equal(count.cpu().numpy(), exp_rt)
    np.testing.assert_array_equal(top50.cpu().numpy(), exp_top50)
    if not ties:                                     # without ties this is exactly the reference's order
        ri, t1 = O.i2t_ranks(S); rt, t50 = O.t2i_ranks(S)
        np.testing.assert_array_equal(rank.cpu().numpy(), ri)
        np.testing.assert_array_equal(top50.cpu().numpy(), t50)


def test_rank_kernels_reproduce_reference_golden():
    from aladin_b200 import ranking
    g = load_golden("retrieval")
    Sd = torch.from_numpy(g["S_full"]).cuda()
    rank, top1 = ranking.rank_rows(Sd)
    count, top50 = ranking.t2i_rank_topk(Sd, 50)
    np.testing.assert_array_equal(rank.cpu().numpy(), g["ranks_i2t"])
    np.testing.assert_array_equal(top1.cpu().numpy(), g["top1"])
    np.testing.assert_array_equal(count.cpu().numpy(), g["ranks_t2i"])
    np.testing.assert_array_equal(top50.cpu().numpy(), g["top50"])


def test_sharded_ranking_equals_single_shard():
    """Image-block shards + gt exchange + count sum + candidate merge == one shard."""
    from aladin_b200 import ranking
    r = np.random.RandomState(5)
    Ni, Nc, G, k = 96, 480, 4, 50
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    Sd = torch.from_numpy(S).cuda()
    full_rank, full_top1 = ranking.rank_rows(Sd)
    full_count, full_topk = ranking.t2i_rank_topk(Sd, k)
    per = Ni // G
    gt = torch.zeros(Nc, device="cuda")
    shards = [Sd[g * per:(g + 1) * per].contiguous() for g in range(G)]
    for g in range(G):
        ranking.col_gt(shards[g], gt, 5, g * per)                # stands in for the all-gather
    counts = sum(ranking.col_count(shards[g], gt, 5, g * per) for g in range(G))
    cands = [ranking.col_topk(shards[g], k, g * per, splits=2 if g % 2 else None) for g in range(G)]
    cs = torch.cat([c[0] for c in cands]); ci = torch.cat([c[1] for c in cands])
    _, merged = ranking.topk_merge(cs, ci)
    ranks = torch.cat([ranking.rank_rows(shards[g], 5, g * per)[0] for g in range(G)])
    assert torch.equal(ranks, full_rank)
    assert torch.equal(counts.int(), full_count)
    assert torch.equal(merged, full_topk)


def _expected_topk(S, k, img_off=0):
    """numpy restatement of the documented order: stable argsort reversed (score desc, index desc on ties)."""
    Ni, Nc = S.shape
    out = np.full((Nc, k), -1, np.int64)
    for c in range(Nc):
        inds = _stable_desc_order(S[:, c])[:k]
        out[c, :inds.size] = inds + img_off
    return out


@pytest.mark.parametrize("Ni,Nc,k,kind", [
    (700, 300, 50, "random"),         # threshold path (Ni > 256 candidates capacity), several row groups
    (2600, 200, 50, "random"),        # many rows per group
    (625, 1000, 50, "ties"),          # heavy exact ties inside the top-k: still ranked by index
    (900, 96, 50, "constant"),        # every column constant -> candidate lists overflow -> heap fallback
    (900, 130, 50, "mixed"),          # constant, -inf-padded and random columns side by side
    (200, 77, 50, "random"),          # Ni <= capacity: no threshold, everything is a candidate
    (30, 40, 50, "random"),           # fewer images than k: tail filled with -1
    (1500, 64, 100, "random"),        # the two-stage shortlist size
    (3000, 33, 200, "random"),        # k above the select limit: heaps over all columns
])
def test_col_topk_select_matches_stable_argsort(Ni, Nc, k, kind):
    from aladin_b200 import ranking
    r = np.random.RandomState(Ni + Nc + k)
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    if kind == "ties":
        S = np.round(S * 2) / 2
    elif kind == "constant":
        S[:] = r.standard_normal(Nc).astype(np.float32)[None, :]
        S[:, ::3] = 0.0
    elif kind == "mixed":
        S[:, 0::4] = 0.0                                            # all-zero captions (no scored words)
        keep = r.rand(Ni, Nc) < 0.05
        S[:, 1::4] = np.where(keep[:, 1::4], S[:, 1::4], -np.inf)   # two-stage style: -inf outside the shortlist
        S[:, 2::4] = np.where(S[:, 2::4] > 0, S[:, 2::4], -0.0)     # signed zeros compare equal
    Sd = torch.from_numpy(S).cuda()
    cs, ci = ranking.col_topk(Sd, k, img_off=7)
    assert cs.shape == (1, Nc, k)
    exp = _expected_topk(S, k, img_off=7)
    got = ci[0].cpu().numpy()
    np.testing.assert_array_equal(got, exp)
    valid = exp >= 0
    np.testing.assert_array_equal(cs[0].cpu().numpy()[valid], S[(exp - 7)[valid], np.nonzero(valid)[0]])
    # and the heap kernel (explicit row slices + merge) agrees entry for entry
    hs, hi = ranking.topk_merge(*ranking.col_topk(Sd, k, img_off=7, splits=3))
    np.testing.assert_array_equal(hi.cpu().numpy(), exp)


def test_col_topk_select_strided_view():
    """Column slices of a wider matrix (ldS > Nc), as rank_both_directions passes them."""
    from aladin_b200 import ranking
    r = np.random.RandomState(3)
    S = r.standard_normal((640, 500)).astype(np.float32)
    Sd = torch.from_numpy(S).cuda()
    _, ci = ranking.col_topk(Sd[:, :333], 50)
    np.testing.assert_array_equal(ci[0].cpu().numpy(), _expected_topk(S[:, :333], 50))


@pytest.mark.parametrize("Ni,Nc,q_rows,q_cols,k,img_off,kind", [
    (1000, 5000, 1000, 5000, 50, 0, "random"),      # COCO-1k: 4 column CTAs (the last one partly filled) x 125 row groups
    (333, 1664, 333, 1664, 50, 0, "ties"),          # heavy exact ties in both directions
    (700, 4096, 500, 2400, 50, 7, "given"),         # query sub-ranges, an image offset, ground truth from outside
    (640, 2052, 640, 2052, 50, 0, "masked"),        # -inf scores (shortlist-masked matrices): exact-comparison branches
    (900, 96, 900, 96, 50, 0, "constant"),          # constant columns: candidate overflow -> heap path
    (2500, 1024, 2500, 1024, 100, 0, "random"),     # the two-stage shortlist size, many rows per group
    (700, 333, 700, 333, 50, 0, "random"),          # row stride not a multiple of four: the one-purpose kernels
    (100, 500, 100, 500, 50, 0, "random"),          # too small for the threshold select: the one-purpose kernels
    (300, 1500, 300, 1500, 50, 0, "nocount"),       # multi-GPU first sweep: no t2i counts
])
def test_rank_fused_equals_one_purpose_kernels(Ni, Nc, q_rows, q_cols, k, img_off, kind):
    """alad_rank_fused (two sweeps of S) against rank_rows + col_gt + col_count + col_topk (four) on the same block:
    every output identical, and the documented total order (numpy stable argsort reversed) on sampled queries."""
    from aladin_b200 import ranking
    r = np.random.RandomState(Ni + Nc)
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    cols = np.arange(Nc)
    own = (cols // 5 - img_off >= 0) & (cols // 5 - img_off < Ni)
    S[(cols // 5 - img_off)[own], cols[own]] += 1.5
    if kind == "ties":
        S = np.round(S * 2) / 2
    if kind == "masked":
        S[r.rand(Ni, Nc) < 0.7] = -np.inf
        S[17] = -np.inf
        S[:, 23] = -np.inf
    if kind == "constant":
        S[:] = 0.25
    Sd = torch.from_numpy(S).cuda()
    gt = None
    if kind == "given":
        g = r.standard_normal(q_cols).astype(np.float32)
        g[::2] = S[(np.arange(q_cols) * 7) % Ni, np.arange(q_cols)][::2]        # exact ties with entries of the block
        gt = torch.from_numpy(g).cuda()
    rank, top1, cnt, ts, ti = ranking.rank_fused(Sd, k, img_off, q_rows, q_cols, gt=gt, count=kind != "nocount")
    rank0, top10 = ranking.rank_rows(Sd[:q_rows], 5, img_off)
    Sq = Sd[:, :q_cols]
    if gt is None:
        gt = torch.zeros(q_cols, dtype=torch.float32, device="cuda")
        ranking.col_gt(Sq, gt, 5, img_off)
    cnt0 = ranking.col_count(Sq, gt, 5, img_off)
    cs0, ci0 = ranking.col_topk(Sq, k, img_off)
    assert torch.equal(rank, rank0) and torch.equal(top1, top10)
    if kind != "nocount":
        assert torch.equal(cnt, cnt0)
    else:
        assert cnt is None
    assert torch.equal(ti, ci0[0]) and torch.equal(ts, cs0[0])
    rank, top1, ti = rank.cpu().numpy(), top1.cpu().numpy(), ti.cpu().numpy()
    for i in range(0, q_rows, 41):
        inds = _stable_desc_order(S[i])
        assert top1[i] == inds[0]
        g0 = 5 * (img_off + i)
        if g0 < Nc:
            pos = np.empty(Nc, np.int64); pos[inds] = np.arange(Nc)
            assert rank[i] == pos[g0:g0 + 5].min()
        else:
            assert rank[i] == Nc
    kk = min(k, Ni)
    for c in range(0, q_cols, 53):
        np.testing.assert_array_equal(ti[c, :kk], _stable_desc_order(S[:, c])[:kk] + img_off)


def test_rank_fused_reproduces_reference_golden():
    from aladin_b200 import ranking
    g = load_golden("retrieval")
    Sd = torch.from_numpy(g["S_full"]).cuda()
    rank, top1, count, _, top50 = ranking.rank_fused(Sd, 50)
    np.testing.assert_array_equal(rank.cpu().numpy(), g["ranks_i2t"])
    np.testing.assert_array_equal(top1.cpu().numpy(), g["top1"])
    np.testing.assert_array_equal(count.cpu().numpy(), g["ranks_t2i"])
    np.testing.assert_array_equal(top50.cpu().numpy(), g["top50"])
