"""Concurrent pinned host->device bandwidth of all ranks of one box, before and after binding every process to the CPUs
next to its GPU (NVML affinity) and re-allocating the pinned source there.  Run under torchrun; rank 0 prints one JSON line."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def measure(src, dst, iters=6):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return src.numel() * src.element_size() * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 256 << 20
    dst = torch.empty(n, dtype=torch.float32, device="cuda")
    out = {"world": world}
    src = torch.empty(n, dtype=torch.float32).pin_memory()
    src.fill_(1.0)
    measure(src, dst, 2)
    a = measure(src, dst)
    info = {}
    try:
        import hostbind
        info = hostbind.bind_to_gpu(local)
    except Exception as e:                      # diagnostics only
        info = {"error": repr(e)}
    del src
    src = torch.empty(n, dtype=torch.float32).pin_memory()
    src.fill_(1.0)
    measure(src, dst, 2)
    b = measure(src, dst)
    rows = [None] * world
    dist.all_gather_object(rows, {"rank": rank, "gbs_unbound": round(a, 2), "gbs_bound": round(b, 2), "bind": info})
    if rank == 0:
        out["ranks"] = rows
        out["sum_unbound"] = round(sum(r["gbs_unbound"] for r in rows), 1)
        out["sum_bound"] = round(sum(r["gbs_bound"] for r in rows), 1)
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
