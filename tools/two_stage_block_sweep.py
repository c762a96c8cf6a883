"""Sweep of the L2 image-block size of the pair-list pass (two_stage.L2_BLOCK_BYTES) at COCO-5k shape, K = 100."""
import json
import sys

sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
import two_stage_probe as P  # noqa: E402
from aladin_b200 import two_stage  # noqa: E402

for mb in [int(x) for x in (sys.argv[1:] or ["24", "32", "48", "64", "96", "400"])]:
    two_stage.L2_BLOCK_BYTES = mb << 20
    r = P.measure(5000, 25000, 100, world=1, steps=3, warmup=2)
    print(json.dumps({"l2_block_mb": mb, "ms_per_call": round(r["ms_per_call"], 3), "pair_kernel_ms": round(r["pair_kernel_ms"], 3),
                      "tiles": r["pair_tiles_rank0"]}), flush=True)
