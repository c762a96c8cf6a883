import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import two_stage_probe as P
from aladin_b200 import two_stage
for flag in (True, False, True, False):
    two_stage.WORD_BOX = flag
    r = P.measure(5000, 25000, 100, world=1, steps=4, warmup=2)
    print(json.dumps({"word_box": flag, "ms_per_call": round(r["ms_per_call"], 3), "pair_kernel_ms": round(r["pair_kernel_ms"], 3), "r1": r["recall_at_1"]}), flush=True)
