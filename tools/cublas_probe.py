"""Same-box context for the roofline: cuBLAS bf16 GEMM, sustained, the way MEASURED_PEAKS.json
was taken (torch.matmul 8192^3 back to back for a few seconds, CUDA events).  Measurement
probe only -- nothing in the product path calls cuBLAS."""
import json
import subprocess
import sys
import time

import torch


def sustained_bf16_tflops(seconds=3.0, n=8192):
    a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(5):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    # calibrate, then queue the whole run without host round trips (a busy host must not starve the GPU)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    iters = max(20, int(seconds * 1e3 / (e0.elapsed_time(e1) / 10)))
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b, out=c)
    e1.record()
    time.sleep(seconds * 0.7)
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"cublas_bf16_tflops_sustained": 2.0 * n ** 3 * iters / (ms * 1e-3) / 1e12, "seconds": ms * 1e-3,
            "clocks_sm_power_at_end": clk}


if __name__ == "__main__":
    print(json.dumps(sustained_bf16_tflops(float(sys.argv[1]) if len(sys.argv) > 1 else 3.0)))
