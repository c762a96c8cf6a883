"""Interleaved in-process sweep of the scoring kernel's L2 settings (ALAD_L2_BLOCK_MB x ALAD_L2_HINTS)
at COCO-5k shape: the operands are packed once, every configuration is launched in rotation
(ABCABC...) so that thermal / power-cap drift hits all of them alike; per launch: CUDA-event time,
SM clock and board power sampled by nvidia-smi while the kernel runs.
usage: python tools/sweep_tile_order.py [rounds] [Ni] [Nc] [fp32]"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from aladin_b200 import _cabi, scoring, synth  # noqa: E402
from aladin_b200.tiling import build_region_tiles  # noqa: E402


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for line in self.p.stdout:
            try:
                c, w = line.split(",")
                self.rows.append((time.perf_counter(), float(c), float(w)))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [(c, w) for t, c, w in self.rows if t0 <= t <= t1]
        if not r:
            return None, None
        return float(np.median([c for c, _ in r])), float(np.median([w for _, w in r]))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    Ni = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    Nc = int(sys.argv[3]) if len(sys.argv) > 3 else 25000
    split = len(sys.argv) > 4 and sys.argv[4] == "fp32"       # 3x split-precision operands (Kp = 3 * d)
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, 34, 50, 1024)
    R, W, nr, nw, clamp = scoring.scored_counts(images.shape, captions.shape, im_len, s_len)
    words = scoring.pack_tokens(captions, nw, slot0=1, mode=1 if split else 0, want_row_item=True)
    regions = scoring.pack_tokens(images, nr, slot0=1, mode=2 if split else 0)
    _, table, _ = build_region_tiles(nr, clamp)
    tiles_dev = scoring._to_dev(table.view(np.int32).reshape(-1), images.device)
    del images, captions
    S = torch.empty((Ni, Nc), dtype=torch.float32, device="cuda")
    configs = [(mb, h) for mb in (8, 12, 16, 24, 30, 45) for h in (0,)] + [(16, 1), (30, 1), (16, 3), (30, 2)]
    if os.environ.get("SWEEP_CONFIGS"):
        # "block:hints[:skew[:group]]" -- skew / group = ALAD_M_SKEW / ALAD_M_SKEW_GROUP (skewed M sweep: an experiment
        # that was measured and removed from the kernel again, profiles/r01_tile_order_sweep.md section 9)
        configs = [tuple(int(x) for x in c.split(":")) for c in os.environ["SWEEP_CONFIGS"].split(",")]
    smi = Smi()
    flops = 2.0 * 34 * 50 * 1024 * Ni * Nc * (3 if split else 1)       # issued FLOP in split mode
    res = {c: [] for c in configs}
    # warm-up
    for _ in range(3):
        scoring.mrsw_scores_packed(words, regions, tiles_dev, len(table), Ni, Nc, out=S)
    torch.cuda.synchronize()
    for r in range(rounds):
        for cfg in configs:
            mb, h = cfg[0], cfg[1]
            os.environ["ALAD_M_SKEW"] = str(cfg[2] if len(cfg) > 2 else 0)
            os.environ["ALAD_M_SKEW_GROUP"] = str(cfg[3] if len(cfg) > 3 else 1)
            os.environ["ALAD_L2_BLOCK_MB"] = str(mb)
            if mb >= 1000:                                   # 1000 + n: block of n tiles
                os.environ["ALAD_N_BLOCK"] = str(mb - 1000)
            else:
                os.environ.pop("ALAD_N_BLOCK", None)
            os.environ["ALAD_L2_HINTS"] = str(h & 3)
            os.environ["ALAD_L2_PREFETCH"] = "1" if h & 4 else "0"     # hints bit 2 = word-row L2 prefetch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(2):
                scoring.mrsw_scores_packed(words, regions, tiles_dev, len(table), Ni, Nc, out=S)
            e1.record()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ms = e0.elapsed_time(e1) / 2
            clk, pw = smi.window(t0 + 0.15, t1)
            res[cfg].append((ms, clk, pw))
    smi.p.terminate()
    out = []
    for cfg, v in res.items():
        mb, h = cfg[0], cfg[1]
        ms = float(np.median([x[0] for x in v]))
        clk = float(np.median([x[1] for x in v if x[1] is not None] or [0]))
        pw = float(np.median([x[2] for x in v if x[2] is not None] or [0]))
        out.append({"l2_block_mb": mb, "hints": h, "skew": list(cfg[2:]), "ms": round(ms, 2), "tflops": round(flops / ms / 1e9, 1), "sm_mhz": clk,
                    "power_w": pw, "all_ms": [round(x[0], 1) for x in v]})
        print(out[-1])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
