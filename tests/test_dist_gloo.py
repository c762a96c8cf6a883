"""CPU, world_size 2, gloo: the multi-GPU exchange protocol of the sharded retrieval
(retrieval.rank_both_directions: ONE all-gather of ground-truth scores + top-k candidates + i2t results + shard
timings, merge, count all-reduce) gives the single-shard answer, with equal and with speed-weighted image blocks.  The per-shard ranking ops are supplied by the
oracle here (tests may use it); on the GPU box the same protocol runs over NCCL with the CUDA ops."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import alad_oracle as O


class OracleOps:
    """numpy stand-ins with the semantics of aladin_b200.ranking (score desc, index desc on ties)."""

    @staticmethod
    def _order(v):
        return np.argsort(v, kind="stable")[::-1]

    @classmethod
    def rank_rows(cls, S, group=5, img_off=0):
        S = S.numpy()
        rank, top1 = np.zeros(len(S), np.int32), np.zeros(len(S), np.int32)
        for i in range(len(S)):
            inds = cls._order(S[i])
            pos = np.empty(len(inds), np.int64)
            pos[inds] = np.arange(len(inds))
            g0 = group * (img_off + i)
            rank[i], top1[i] = pos[g0:g0 + group].min(), inds[0]
        return torch.from_numpy(rank), torch.from_numpy(top1)

    @staticmethod
    def col_gt(S, gt, group=5, img_off=0):
        for c in range(S.shape[1]):
            i = c // group - img_off
            if 0 <= i < S.shape[0]:
                gt[c] = S[i, c]
        return gt

    @staticmethod
    def col_count(S, gt, group=5, img_off=0):
        S, g = S.numpy(), gt.numpy()
        idx = img_off + np.arange(S.shape[0])[:, None]
        gimg = (np.arange(S.shape[1]) // group)[None, :]
        ahead = (S > g[None]) | ((S == g[None]) & (idx > gimg))
        return torch.from_numpy(ahead.sum(0).astype(np.int32))

    @classmethod
    def col_topk(cls, S, k, img_off=0, splits=1):
        S = S.numpy()
        cs = np.full((1, S.shape[1], k), -np.inf, np.float32)
        ci = np.full((1, S.shape[1], k), -1, np.int32)
        for c in range(S.shape[1]):
            inds = cls._order(S[:, c])[:k]
            cs[0, c, :len(inds)], ci[0, c, :len(inds)] = S[inds, c], inds + img_off
        return torch.from_numpy(cs), torch.from_numpy(ci)

    @staticmethod
    def topk_merge(cs, ci):
        cs, ci = cs.numpy(), ci.numpy()
        P, Nc, k = cs.shape
        os_, oi = np.zeros((Nc, k), np.float32), np.zeros((Nc, k), np.int32)
        for c in range(Nc):
            s, i = cs[:, c].reshape(-1), ci[:, c].reshape(-1)
            order = np.lexsort((-i, -s))[:k]                       # score desc, then index desc
            os_[c], oi[c] = s[order], i[order]
        return torch.from_numpy(os_), torch.from_numpy(oi)


def _worker(rank, world, port, S_full, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aladin_b200 import retrieval
        Ni = S_full.shape[0]
        lo, hi = retrieval.shard_bounds(Ni, world, rank)
        res = retrieval.rank_both_directions(S_full[lo:hi].clone(), Ni, img_off=lo, n_images_total=Ni, k=50,
                                             group=dist.group.WORLD, ops=OracleOps)
        # speed-weighted blocks: rank 1 measured 1.5x faster -> all ranks derive the same uneven bounds
        retrieval.balancer.update(world, [30, 31], [3.0, 2.0])
        bounds = retrieval.balancer.all_bounds(Ni, world)
        lo, hi = bounds[rank]
        res2 = retrieval.rank_both_directions(S_full[lo:hi].clone(), Ni, img_off=lo, n_images_total=Ni, k=50,
                                              group=dist.group.WORLD, ops=OracleOps, bounds=bounds)
        if rank == 0:
            out.put([np.asarray(r) for r in res] + [np.asarray(r) for r in res2] + [np.asarray(bounds)])
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_ranking_protocol_world2_gloo():
    r = np.random.RandomState(3)
    Ni, Nc = 61, 305                                               # odd size: uneven image blocks
    S = r.standard_normal((Ni, Nc)).astype(np.float32)
    S[np.arange(Nc) // 5, np.arange(Nc)] += 1.0
    S_t = torch.from_numpy(S)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, S_t, out)) for rk in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    ranks_i2t, top1, ranks_t2i, top50 = got[:4]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ri, t1 = O.i2t_ranks(S)
    rt, t50 = O.t2i_ranks(S, 50)
    np.testing.assert_array_equal(ranks_i2t, ri)
    np.testing.assert_array_equal(top1, t1)
    np.testing.assert_array_equal(ranks_t2i, rt)
    np.testing.assert_array_equal(top50, t50)
    bounds = got[8]
    assert bounds[0][1] - bounds[0][0] < bounds[1][1] - bounds[1][0] and bounds[0][0] == 0 and bounds[1][1] == Ni
    for a, b in zip(got[4:8], (ri, t1, rt, t50)):
        np.testing.assert_array_equal(a, b)


def test_shard_balancer_tracks_measured_speed():
    from aladin_b200 import retrieval
    b = retrieval.ShardBalancer()
    assert b.all_bounds(5000, 4) == [retrieval.shard_bounds(5000, 4, r) for r in range(4)]
    for _ in range(6):                                   # rank 2 is 10 % slower than the others
        spans = b.all_bounds(5000, 4)
        n = [hi - lo for lo, hi in spans]
        b.update(4, n, [n[0] / 1.0, n[1] / 1.0, n[2] / 0.9, n[3] / 1.0])
    spans = b.all_bounds(5000, 4)
    n = np.array([hi - lo for lo, hi in spans], np.float64)
    t = n / np.array([1.0, 1.0, 0.9, 1.0])
    assert spans[0][0] == 0 and spans[-1][1] == 5000 and all(a[1] == c[0] for a, c in zip(spans, spans[1:]))
    assert t.max() / t.min() < 1.01                      # equal finishing times
    b.update(4, n, [0, 1, 1, 1])                         # a step without a measurement changes nothing
    assert b.all_bounds(5000, 4) == spans


def test_shard_bounds_cover_everything():
    from aladin_b200 import retrieval
    for n in (0, 1, 7, 5000):
        for w in (1, 2, 4, 8):
            spans = [retrieval.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _emulated_worker(rank, world, port, S_full, lib_path, k, out):
    """One rank of the sharded ranking with the REAL ranking entry points (alad_rank_fused without counts, alad_col_gt,
    alad_col_count, alad_topk_merge) from the host-thread emulation of csrc/rank.cu; the exchange runs over gloo."""
    import ctypes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aladin_b200 import _cabi, ranking, retrieval
        lib = ctypes.CDLL(lib_path)
        for name, (res, args) in _cabi.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _cabi._lib = lib
        _cabi.stream_ptr = lambda: None
        torch.Tensor.is_cuda = property(lambda self: True)          # "device" tensors are host tensors in this process
        Ni = S_full.shape[0]
        lo, hi = retrieval.shard_bounds(Ni, world, rank)
        block = torch.from_numpy(np.ascontiguousarray(S_full.numpy()[lo:hi]))
        res = retrieval.rank_device(block, Ni, img_off=lo, n_images_total=Ni, k=k, group=dist.group.WORLD, ops=ranking)
        if rank == 0:
            out.put([np.asarray(r) for r in res[:5]])
    finally:
        dist.destroy_process_group()


def test_sharded_ranking_world2_gloo_with_the_emulated_kernels():
    """The N > 1 composition the GPU box runs (fused sweep without counts -> one all-gather -> counts against the
    gathered ground truth -> all-reduce), on the real kernels' sources: equals the stable-argsort order of the whole matrix."""
    import shutil
    import sys
    import pytest
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
    import build_emu
    lib_path = build_emu.build_library()
    r = np.random.RandomState(11)
    Ni, Nc, k = 540, 200, 10                                       # 270 rows per rank: the threshold-select / fused-sweep path
    S = np.round(r.standard_normal((Ni, Nc)).astype(np.float32) * 4) / 4          # exact ties
    S[np.arange(Nc) // 5, np.arange(Nc)] += 1.0
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emulated_worker, args=(rk, 2, port, torch.from_numpy(S), lib_path, k, out)) for rk in range(2)]
    for p in procs:
        p.start()
    rank_i, top1, count, ts, ti = out.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    order = lambda v: np.argsort(v, kind="stable")[::-1]
    for i in range(Ni):
        inds = order(S[i])
        assert top1[i] == inds[0]
        if 5 * i < Nc:
            pos = np.empty(Nc, np.int64)
            pos[inds] = np.arange(Nc)
            assert rank_i[i] == pos[5 * i:5 * i + 5].min()
        else:
            assert rank_i[i] == Nc
    for c in range(Nc):
        inds = order(S[:, c])
        assert count[c] == np.where(inds == c // 5)[0][0]
        np.testing.assert_array_equal(ti[c], inds[:k])
        np.testing.assert_array_equal(ts[c], S[inds[:k], c])
