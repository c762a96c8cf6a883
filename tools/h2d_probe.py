"""Does a small pageable cudaMemcpyAsync (what torch's .to(non_blocking=True) and alad_scores_fused issue for the
per-call metadata) wait for work already queued on the stream?  Queue ~50 ms of GEMMs, then time the host call."""
import time

import numpy as np
import torch

a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
meta = np.arange(4096, dtype=np.int64)
for _ in range(3):
    (a @ a)
torch.cuda.synchronize()
for label, busy in (("idle stream", False), ("busy stream", True)):
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        if busy:
            for _ in range(40):
                (a @ a)
        t0 = time.perf_counter()
        torch.from_numpy(meta).to("cuda", non_blocking=True)
        ts.append(1e6 * (time.perf_counter() - t0))
    torch.cuda.synchronize()
    print(f"{label}: pageable 32 KB H2D host time median {sorted(ts)[5]:.1f} us (min {min(ts):.1f}, max {max(ts):.1f})")
