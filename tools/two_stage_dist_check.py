"""Two-stage retrieval under torchrun (N GPUs of one box): the sharded result must equal the single-GPU one.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/two_stage_dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aladin_b200 import synth, two_stage
    ok = True
    for seed, Ni, d, K, precision in ((7, 203, 128, 20, "fp32"), (8, 640, 256, 100, "bf16")):
        images, captions, il, cl = synth.eval_containers(seed, Ni, 53, d, max_regions=34, max_words=50, alpha=0.25)
        ti, tc = torch.from_numpy(images).pin_memory(), torch.from_numpy(captions).pin_memory()
        _, (ri, rt) = two_stage.two_stage_retrieval(ti, tc, il, cl, shortlist=K, precision=precision, return_ranks=True)
        _, (ri1, rt1) = two_stage.two_stage_retrieval(ti, tc, il, cl, shortlist=K, precision=precision, return_ranks=True,
                                                      group=None)
        # the sharded pass scores a caption in different tiles than the unsharded one only in the ORDER of its image
        # slots; every score is one ordered sum over the caption's words in both -> bit-identical ranks
        same = np.array_equal(ri, ri1) and np.array_equal(rt, rt1)
        print(f"[rank {rank}/{world}] Ni={Ni} K={K} {precision}: sharded == unsharded: {same} "
              f"(i2t R@1 {100 * np.mean(ri < 1):.1f}, t2i R@1 {100 * np.mean(rt < 1):.1f})", flush=True)
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("two_stage_dist_check ok")


if __name__ == "__main__":
    main()
