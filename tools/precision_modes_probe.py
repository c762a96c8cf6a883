import json, sys, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import bench_also
for prec in ("tf32", "fp32", "bf16"):
    ms, k = bench_also.retrieval_step(5000, 25000, prec, steps=2, warmup=2)
    print(json.dumps({"precision": prec, "ms_per_step": round(ms,2), "kernel_ms": round(k,2)}), flush=True)
