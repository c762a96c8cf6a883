"""End-to-end evaluation step from PAGEABLE host tensors (what the reference's encode_data returns) for several numbers of
staging threads (alad_h2d_2d_staged), next to the raw staged-upload bandwidth.  One GPU.  Prints JSON lines."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aladin_b200 import evaluation, loss as L, retrieval, synth  # noqa: E402

Ni, Nc, regions, words, d = bench.WORKLOADS["coco5k"]
images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
imgs_p, caps_p = bench.host_layout(images, captions, pinned=False)
del images, captions
img_lens5 = [l for l in im_len for _ in range(5)]
scorer = L.AlignmentContrastiveLoss(aggregation="MrSw")
fn = evaluation.fused_sim_function(scorer)


def step():
    evaluation.clear_cache()
    evaluation.i2t(imgs_p, caps_p, img_lens5, s_len, sim_function=fn, cap_batches=5)
    evaluation.t2i(imgs_p, caps_p, img_lens5, s_len, sim_function=fn, im_batches=5)


for n in [int(x) for x in (sys.argv[1:] or ["1", "4", "8", "16"])]:
    os.environ["ALAD_H2D_THREADS"] = str(n)
    # raw upload bandwidth of 4096 captions (870 MB)
    dst = torch.empty((4096, 51, d), dtype=torch.float32, device="cuda")
    retrieval._upload_rows(caps_p, 0, 1, 4096, 51, out=dst)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(3):
        retrieval._upload_rows(caps_p, 4096 * k, 1, 4096, 51, out=dst)
    torch.cuda.synchronize()
    gbs = 3 * 4096 * 51 * d * 4 / (time.perf_counter() - t0) / 1e9
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    print(json.dumps({"staging_threads": n, "staged_upload_GBs": round(gbs, 1), "e2e_pageable_ms": round((time.perf_counter() - t0) / 2 * 1e3, 1)}), flush=True)
