"""B x B loss kernels at B = 8192 inside a cudaProfilerStart/Stop range, for the ncu launch list of the
HBM-bound kernels (profiles/*_hbm_kernels.md):
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/loss_launches.csv python tools/loss_probe.py [B]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from aladin_b200 import loss as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
r = np.random.RandomState(B)
S = torch.tensor(r.standard_normal((B, B)).astype(np.float32), device="cuda")
M = torch.tensor(np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32), device="cuda")
T = S * 2 + 3


def run():
    L.triplet_fwd_bwd(S, 0.2, True)       # hardest negatives (all shipped configs)
    L.triplet_fwd_bwd(S, 0.2, False)      # sum over all violations
    L.listnet_fwd_bwd(T, M)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
