"""Ranking kernels on a COCO-5k-sized score matrix: CUDA-event time per piece, threshold-select top-k vs the
heap kernel (entry-for-entry equality), achieved GB/s against the 4*Ni*Nc bytes of one sweep of S.
usage: python tools/rank_probe.py [Ni] [Nc] [k]      (run under ncu for per-kernel DRAM bytes)"""
import json
import sys

import torch

sys.path.insert(0, ".")
from aladin_b200 import ranking  # noqa: E402


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    Ni = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    Nc = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    g = torch.Generator(device="cuda").manual_seed(1)
    S = torch.randn((Ni, Nc), generator=g, device="cuda") * 3 + 20
    S[torch.arange(Nc, device="cuda") // 5 % Ni, torch.arange(Nc, device="cuda")] += 4.0
    sweep = 4.0 * Ni * Nc
    res = {"Ni": Ni, "Nc": Nc, "k": k, "sweep_bytes": sweep}
    gt = torch.zeros(Nc, device="cuda")
    t, _ = timeit(lambda: ranking.rank_rows(S), iters)
    res["rank_rows_ms"] = t
    t, _ = timeit(lambda: ranking.col_gt(S, gt), iters)
    res["col_gt_ms"] = t
    t, _ = timeit(lambda: ranking.col_count(S, gt), iters)
    res["col_count_ms"] = t
    t_sel, (cs, ci) = timeit(lambda: ranking.col_topk(S, k), iters)
    res["col_topk_select_ms"] = t_sel
    res["col_topk_select_GBs_per_sweep"] = sweep / (t_sel * 1e-3) / 1e9
    t_heap, (hs, hi) = timeit(lambda: ranking.topk_merge(*ranking.col_topk(S, k, splits=8)), max(2, iters // 3))
    res["col_topk_heap8_merge_ms"] = t_heap
    res["select_equals_heap"] = bool(torch.equal(ci[0], hi) and torch.equal(cs[0], hs))
    for name in ("rank_rows_ms", "col_count_ms"):
        res[name.replace("_ms", "_GBs")] = sweep / (res[name] * 1e-3) / 1e9
    print(json.dumps(res, indent=1))
    assert res["select_equals_heap"]


if __name__ == "__main__":
    main()
