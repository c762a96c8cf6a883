"""Where the non-kernel time of one bench step goes: host wall-clock and CUDA-event times per phase
of the device-resident step for one shard (world / rank emulated on one GPU, no collectives).
usage: python tools/step_breakdown.py [world] [rank] [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from aladin_b200 import retrieval, synth  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    Ni, Nc = 5000, 25000
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, 34, 50, 1024)
    torch.cuda.synchronize()
    for it in range(steps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision="bf16", world=world, rank=rank)
        t1 = time.perf_counter()
        S = gal.scores()
        e[1].record()
        t2 = time.perf_counter()
        out = retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=None)
        e[2].record()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it >= 2:
            print(f"step {it}: host ms: init {1e3 * (t1 - t0):.2f}  scores() launch {1e3 * (t2 - t1):.2f}  rank+D2H {1e3 * (t3 - t2):.2f}  "
                  f"total {1e3 * (t3 - t0):.2f} | device ms: pack+scores {e[0].elapsed_time(e[1]):.2f}  rank {e[1].elapsed_time(e[2]):.2f}")


if __name__ == "__main__":
    main()
