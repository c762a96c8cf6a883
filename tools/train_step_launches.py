"""One fused training step (forward + backward) at B = 512 and B = 128 inside a cudaProfilerStart/Stop range, for an ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/train_step_launches.py"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import bench_train_step as TS  # noqa: E402

steps = {}
for B in (512, 128):
    fwd, fwd_bwd, il, cl = TS.make_step(B, "bf16", fused=True)
    for _ in range(3):
        fwd_bwd()
    steps[B] = fwd_bwd
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for B in (512, 128):
    steps[B]()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
