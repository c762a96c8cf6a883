#!/bin/bash
# sweep of the region-tile block size kept hot in L2 (ALAD_L2_BLOCK_MB), same box, same process setup
for mb in 64 15 8 30 100 64 15; do
  ALAD_L2_BLOCK_MB=$mb timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('L2_BLOCK_MB=$mb ms/step', round(d['ms_per_step'],2), 'kernel ms', round(r['avg_launch_ms'],2), 'TF', round(r['achieved'],1), 'clk', d['clocks']['sm_mhz'])
"
done
