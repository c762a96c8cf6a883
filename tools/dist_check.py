"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
Every rank runs the public drop-ins (aladin_b200.evaluation.i2t / t2i) on the same seeded host
containers with the gallery images sharded across the ranks (NCCL exchange of ground-truth
scores, counts and top-50 candidates), then recomputes everything unsharded on its own GPU and
requires identical ranks / top-1 / top-50 / metrics."""
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
    from aladin_b200 import evaluation, loss as L, retrieval, synth
    ok = True
    # the host-resident sharded path replicates the packed captions through peer windows (copy engines) or NCCL
    for seed, Ni, d, mr, mw, precision, exchange in ((5, 203, 128, 36, 30, "fp32", "peer"), (6, 640, 256, 34, 50, "bf16", "peer"),
                                                     (6, 640, 256, 34, 50, "bf16", "nccl"), (8, 900, 64, 34, 50, "bf16", "peer")):
        retrieval.EXCHANGE = exchange
        images, captions, img_lens, cap_lens = synth.eval_containers(seed, Ni, 71, d, max_regions=mr, max_words=mw)
        ti, tc = torch.from_numpy(images).pin_memory(), torch.from_numpy(captions).pin_memory()
        crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
        crit.precision = precision
        evaluation.clear_cache()
        mi, (ri, t1) = evaluation.i2t(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=crit, cap_batches=5)
        mt, (rt, t50) = evaluation.t2i(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=crit, im_batches=5)
        # unsharded on this GPU
        gal = retrieval.AlignmentGallery(ti, tc, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5,
                                         precision=precision, world=1, rank=0)
        S = gal.scores()
        ri1, t11, rt1, t501 = retrieval.rank_both_directions(S, Ni, k=50, group=None)
        same = (np.array_equal(ri, ri1) and np.array_equal(t1, t11) and np.array_equal(rt, rt1) and np.array_equal(t50, t501)
                and mi[:5] == retrieval.recall_tuple(ri1) and mt[:5] == retrieval.recall_tuple(rt1))
        print(f"[rank {rank}/{world}] Ni={Ni} {precision} exchange={retrieval.EXCHANGE}: sharded == unsharded: {same}; "
              f"i2t R@1 {mi[0]:.1f} t2i R@1 {mt[0]:.1f}", flush=True)
        ok = ok and retrieval.EXCHANGE == exchange           # a silent fall-back to NCCL is a failure here
        ok = ok and same
    # the two exchanges deliver the same packed rows into the same layout: the shard's score block must be bit-identical,
    # and bit-identical to the unsharded block as well (same cuts of every caption's rows, retrieval._phase_plans)
    for seed, Ni, d, mr, mw in ((7, 1300, 64, 20, 30), (9, 2100, 128, 34, 50)):
        images, captions, img_lens, cap_lens = synth.eval_containers(seed, Ni, 71, d, max_regions=mr, max_words=mw)
        ti, tc = torch.from_numpy(images).pin_memory(), torch.from_numpy(captions).pin_memory()
        blocks = {}
        for exchange in ("peer", "nccl", "peer"):
            retrieval.EXCHANGE = exchange
            gal = retrieval.AlignmentGallery(ti, tc, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5, precision="bf16",
                                             world=world, rank=rank, bounds=[retrieval.shard_bounds(Ni, world, r) for r in range(world)])
            S = gal.scores(group=dist.group.WORLD)
            torch.cuda.synchronize()
            ok = ok and retrieval.EXCHANGE == exchange
            blocks.setdefault(exchange, []).append(S)
        one = retrieval.AlignmentGallery(ti, tc, img_lens, cap_lens, n_images=Ni, img_start=0, img_step=5, precision="bf16",
                                         world=1, rank=0).scores()[gal.lo:gal.hi]
        same = all(torch.equal(blocks["peer"][0], x) for x in blocks["peer"][1:] + blocks["nccl"])
        close = bool(torch.equal(blocks["peer"][0], one))        # shares start at their canonical row modulo the 256-row work unit
        print(f"[rank {rank}/{world}] Ni={Ni}: peer == nccl bit for bit: {same}; == unsharded block bit for bit: {close}", flush=True)
        ok = ok and same and close
    # device-resident gallery under the cross-GPU work pool (steal.py): chunks of a rank's block may be scored by
    # another GPU and added into the owner's block over NVLink -- the block must equal the unsharded one bit for bit
    for seed, Ni, d, mr, mw in ((10, 1500, 128, 34, 50), (11, 700, 64, 20, 30)):
        images, captions, img_lens, cap_lens = synth.eval_containers(seed, Ni, 71, d, max_regions=mr, max_words=mw)
        ti, tc = torch.from_numpy(images[0::5].copy()).cuda(), torch.from_numpy(captions).cuda()
        il = img_lens[0::5]
        one = retrieval.AlignmentGallery(ti, tc, il, cap_lens, n_images=Ni, precision="bf16", world=1, rank=0).scores()
        for use_pool in (True, False, True):
            retrieval.POOL = use_pool
            gal = retrieval.AlignmentGallery(ti, tc, il, cap_lens, n_images=Ni, precision="bf16", world=world, rank=rank)
            S = gal.scores(group=dist.group.WORLD)
            out = retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=dist.group.WORLD,
                                                 bounds=gal.bounds)
            torch.cuda.synchronize()
            same = bool(torch.equal(S, one[gal.lo:gal.hi]))
            ref = retrieval.rank_both_directions(one, Ni, k=50, group=None)
            same_ranks = all(np.array_equal(a, b) for a, b in zip(out, ref))
            pool = next(iter(retrieval._pools.values()), None)
            print(f"[rank {rank}/{world}] Ni={Ni} device-resident, pool={use_pool}: block == unsharded bit for bit: {same}; "
                  f"ranks equal: {same_ranks}; chunks own/taken so far: {pool.stats if pool else None}", flush=True)
            ok = ok and same and same_ranks
        retrieval.POOL = True
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    for xc in retrieval._exchanges.values():
        xc.check_error_async(torch.cuda.current_stream())
        torch.cuda.synchronize()
        if int(xc.error_host[0]) != 0:
            print(f"[rank {rank}] peer wait timed out (rank {int(xc.error_host[0]) - 1})", flush=True)
            sys.exit(1)
    retrieval.close_exchanges()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist_check ok")


if __name__ == "__main__":
    main()
