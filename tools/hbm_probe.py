"""HBM-bound kernels of the path, CUDA-event times on one B200 (round 2, second half):
  * ranking at COCO-5k size: the four one-purpose sweeps (rank_rows, col_count, col_topk = group maxima + collect)
    against alad_rank_fused (two sweeps), entry-for-entry equality of every output, GB/s against 4*Ni*Nc per sweep;
  * ListNet and triplet at B = 8192 and B = 512: time, algorithmic GB/s (16*B^2 / 8*B^2 bytes), and the ListNet loss /
    gradient against a torch fp64 restatement of alad/loss.py:427-445 on the same device.
usage: python tools/hbm_probe.py [out.json]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from aladin_b200 import loss as L, ranking  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return max(e0.elapsed_time(e1) / iters, 1e-9), out


def ranking_part(Ni=5000, Nc=25000, k=50):
    torch.manual_seed(1)
    S = torch.randn((Ni, Nc), device="cuda") * 3 + 20
    S[torch.arange(Nc, device="cuda") // 5 % Ni, torch.arange(Nc, device="cuda")] += 4.0
    S[:, 7] = torch.round(S[:, 7])                       # a column and a row full of exact ties
    S[11] = torch.round(S[11])
    sweep = 4.0 * Ni * Nc

    def separate():
        rk, t1 = ranking.rank_rows(S)
        gt = torch.zeros(Nc, dtype=torch.float32, device="cuda")
        ranking.col_gt(S, gt)
        cnt = ranking.col_count(S, gt)
        cs, ci = ranking.col_topk(S, k)
        return rk, t1, cnt, cs[0], ci[0]

    t_sep, a = timeit(separate)
    t_fus, b = timeit(lambda: ranking.rank_fused(S, k))
    t_nc, c = timeit(lambda: ranking.rank_fused(S, k, count=False))
    equal = all(torch.equal(x, y) for x, y in zip(a, b)) and torch.equal(a[0], c[0]) and torch.equal(a[4], c[4])
    pieces = {}
    gt = torch.zeros(Nc, dtype=torch.float32, device="cuda")
    ranking.col_gt(S, gt)
    pieces["rank_rows_ms"], _ = timeit(lambda: ranking.rank_rows(S))
    pieces["col_count_ms"], _ = timeit(lambda: ranking.col_count(S, gt))
    pieces["col_topk_select_ms"], _ = timeit(lambda: ranking.col_topk(S, k))
    res = {"Ni": Ni, "Nc": Nc, "k": k, "sweep_bytes": sweep, "separate_ms": t_sep, "fused_ms": t_fus,
           "fused_without_counts_ms": t_nc, "fused_equals_separate": bool(equal), **pieces,
           "separate_GBs_algorithmic": sweep / (t_sep * 1e-3) / 1e9, "fused_GBs_algorithmic": sweep / (t_fus * 1e-3) / 1e9}
    # masked matrix (two-stage style): -inf everywhere except a shortlist
    S2 = torch.full_like(S, float("-inf"))
    keep = torch.rand((Ni, Nc), device="cuda") < 0.02
    S2[keep] = S[keep]
    a2 = None

    def separate2():
        rk, t1 = ranking.rank_rows(S2)
        gt2 = torch.zeros(Nc, dtype=torch.float32, device="cuda")
        ranking.col_gt(S2, gt2)
        return rk, t1, ranking.col_count(S2, gt2), *[x[0] for x in ranking.col_topk(S2, k)]

    a2 = separate2()
    b2 = ranking.rank_fused(S2, k)
    res["masked_equal"] = bool(all(torch.equal(x, y) for x, y in zip(a2, b2)))
    return res


def listnet_ref(T, M, tau=6.0, eps=1e-10):
    T64, M64 = T.double(), M.double().requires_grad_(True)
    total = 0
    for dim in (0, 1):
        p = torch.softmax(M64 * tau, dim=dim)
        t = torch.softmax(T64, dim=dim)
        total = total + (-(t * torch.log(p + eps)).sum(dim=dim)).mean()
    total.backward()
    return float(total), M64.grad.float()


def loss_part(B):
    r = np.random.RandomState(B)
    S = torch.tensor(r.standard_normal((B, B)).astype(np.float32), device="cuda")
    M = torch.tensor(np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32), device="cuda")
    T = S * 2 + 3
    iters = 20 if B > 2048 else 200
    t_trip, _ = timeit(lambda: L.triplet_fwd_bwd(S, 0.2, True), iters)
    t_list, (loss, dM) = timeit(lambda: L.listnet_fwd_bwd(T, M), iters)
    ref_loss, ref_g = listnet_ref(T, M)
    gerr = float((dM - ref_g).abs().max() / ref_g.abs().max())
    return {"B": B, "triplet_ms": t_trip, "triplet_GBs_algorithmic": 8.0 * B * B / (t_trip * 1e-3) / 1e9,
            "listnet_ms": t_list, "listnet_GBs_algorithmic": 16.0 * B * B / (t_list * 1e-3) / 1e9,
            "listnet_loss": float(loss), "listnet_loss_ref_fp64": ref_loss,
            "listnet_loss_rel_err": abs(float(loss) - ref_loss) / abs(ref_loss), "listnet_grad_max_err_rel_to_max": gerr}


def main():
    out = {"ranking": ranking_part(), "losses": [loss_part(8192), loss_part(512), loss_part(1000)]}
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")
    assert out["ranking"]["fused_equals_separate"] and out["ranking"]["masked_equal"]
    for l in out["losses"]:
        assert l["listnet_loss_rel_err"] < 2e-5 and l["listnet_grad_max_err_rel_to_max"] < 2e-3, l


if __name__ == "__main__":
    main()
