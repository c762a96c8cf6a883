"""cProfile of the device-resident bench step under torchrun (rank 0 prints): where the HOST time of a step goes."""
import cProfile
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD if world > 1 else None
    import bench
    from aladin_b200 import retrieval, synth
    Ni, Nc, regions, words, d = bench.WORKLOADS["coco5k"]
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
    fake_world = int(os.environ.get("FAKE_WORLD", world))       # time an 8-way shard's step on fewer GPUs (no collectives then)

    def step():
        if fake_world != world:
            gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision="bf16", world=fake_world, rank=0)
            S = gal.scores()
            return retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=None, lists_to_host=False)
        gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision="bf16", world=world, rank=rank)
        S = gal.scores()
        return retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=group, bounds=gal.bounds,
                                              lists_to_host=False)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    # host wall time from "GPU idle" to "results on the host", against the device time of the same step
    rows = []
    for _ in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        step()
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        rows.append((1e3 * (t1 - t0), e0.elapsed_time(e1)))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    if rank == 0:
        print("host wall / device ms per step:", [(round(a, 3), round(b, 3)) for a, b in rows])
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
