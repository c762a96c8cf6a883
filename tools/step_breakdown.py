"""Where one bench step of a rank goes when the COCO-5k gallery is sharded `world` ways (one GPU, no collectives):
host wall time (launch queue empty at the start) and device time of  gallery construction (host bookkeeping) ->
pack + scores -> ranking -> results to host.    python tools/step_breakdown.py [world] [workload]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from aladin_b200 import retrieval, synth  # noqa: E402
from bench import WORKLOADS  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Ni, Nc, regions, words, d = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "coco5k"]
images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
torch.cuda.synchronize()


def step(rec):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [time.perf_counter()]
    ev[0].record()
    gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision="bf16", world=world, rank=0)
    t.append(time.perf_counter())
    ev[1].record()
    S = gal.scores()
    t.append(time.perf_counter())
    ev[2].record()
    out = retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=None)
    t.append(time.perf_counter())
    ev[3].record()
    torch.cuda.synchronize()
    if rec is not None:
        rec.append({"host_ms": {"gallery_ctor": 1e3 * (t[1] - t[0]), "scores_enqueue": 1e3 * (t[2] - t[1]),
                                "rank_and_d2h_incl_wait": 1e3 * (t[3] - t[2]), "total": 1e3 * (t[3] - t[0])},
                    "device_ms": {"ctor": ev[0].elapsed_time(ev[1]), "pack_scores": ev[1].elapsed_time(ev[2]),
                                  "rank_d2h": ev[2].elapsed_time(ev[3]), "total": ev[0].elapsed_time(ev[3])}})
    return out


for _ in range(3):
    step(None)
rec = []
for _ in range(5):
    step(rec)
med = lambda xs: sorted(xs)[len(xs) // 2]   # noqa: E731
summary = {k: {kk: round(med([r[k][kk] for r in rec]), 3) for kk in rec[0][k]} for k in rec[0]}
print(json.dumps({"world": world, "rank": 0, "images_of_rank": (Ni + world - 1) // world, "captions": Nc, **summary}))

if len(sys.argv) > 3 and sys.argv[3] == "hostprof":
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        step(None)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
