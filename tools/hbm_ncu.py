"""One launch of each HBM-bound composition inside a cudaProfilerStart/Stop range, for ncu:
    ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/hbm python tools/hbm_ncu.py
(ranking of a COCO-5k-sized score matrix through alad_rank_fused with and without the t2i counts; ListNet and the
triplet loss at B = 8192)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from aladin_b200 import loss as L, ranking  # noqa: E402

Ni, Nc, k, B = 5000, 25000, 50, 8192
torch.manual_seed(1)
S = torch.randn((Ni, Nc), device="cuda") * 3 + 20
S[torch.arange(Nc, device="cuda") // 5 % Ni, torch.arange(Nc, device="cuda")] += 4.0
r = np.random.RandomState(B)
X = torch.tensor(r.standard_normal((B, B)).astype(np.float32), device="cuda")
M = torch.tensor(np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32), device="cuda")
T = X * 2 + 3


def run():
    ranking.rank_fused(S, k)
    ranking.rank_fused(S, k, count=False)
    L.listnet_fwd_bwd(T, M)
    L.triplet_fwd_bwd(X, 0.2, True)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
