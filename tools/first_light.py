"""Bring-up probe for the tcgen05 scoring kernel: correctness diagnostics + a first timing.
Run on the GPU box:  python tools/first_light.py [gemm|mrsw|time]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from aladin_b200 import scoring, synth, tiling  # noqa: E402


def gemm(Ni=500, Nc=300, d=192):
    r = np.random.RandomState(7)
    im = (r.randint(-4, 5, size=(Ni, d)) / 8.0).astype(np.float32)
    s = (r.randint(-4, 5, size=(Nc, d)) / 8.0).astype(np.float32)
    got = scoring.dot_scores(torch.from_numpy(im).cuda(), torch.from_numpy(s).cuda(), precision="bf16")
    torch.cuda.synchronize()
    got = got.cpu().numpy().astype(np.float64)
    ref = im.astype(np.float64) @ s.astype(np.float64).T
    bad = np.argwhere(got != ref)
    print(f"gemm {Ni}x{Nc}x{d}: mismatches {len(bad)} / {got.size}, max abs err {np.abs(got - ref).max():.4g}")
    if len(bad):
        print("first mismatches (region row n, word row m, got, ref):")
        for n, m in bad[:12]:
            print(f"  n={n} m={m} got={got[n, m]:.4f} ref={ref[n, m]:.4f}")
        rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
        print("bad region rows:", rows_bad[:40], "... count", len(rows_bad))
        print("bad word rows:", cols_bad[:40], "... count", len(cols_bad))
    return len(bad) == 0


def timing(Ni=1000, Nc=5000, iters=5):
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc)
    torch.cuda.synchronize()
    for prec in ("bf16", "fp32"):
        for _ in range(2):
            S = scoring.alignment_scores(images, captions, im_len, s_len, precision=prec)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(iters):
            S = scoring.alignment_scores(images, captions, im_len, s_len, precision=prec)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        flops = 2.0 * 34 * 50 * 1024 * Ni * Nc * (3 if prec == "fp32" else 1)
        print(f"alignment_scores {Ni}x{Nc} {prec}: {ms:.3f} ms/iter, {Ni * Nc / ms * 1e3:.3e} pairs/s, "
              f"{flops / ms / 1e9:.1f} TFLOP/s issued")
    print("S stats", float(S.mean()), float(S.max()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    t0 = time.time()
    if what == "gemm":
        ok = gemm(240, 128, 64) and gemm(500, 300, 192) and gemm(7, 1000, 1024)
        sys.exit(0 if ok else 1)
    elif what == "time":
        timing(*(int(a) for a in sys.argv[2:4])) if len(sys.argv) > 3 else timing()
    print(f"done in {time.time() - t0:.1f}s")
