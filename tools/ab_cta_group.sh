#!/bin/bash
# A/B of the two scoring-kernel variants (ALAD_CTA_GROUP=1|2): exactness tests, then timings.
for cg in 2 1; do
  echo "=============== ALAD_CTA_GROUP=$cg"
  ALAD_CTA_GROUP=$cg timeout 300 python -m pytest tests/test_gpu_scoring.py -x -q 2>&1 | tail -4
  ALAD_CTA_GROUP=$cg timeout 200 python tools/first_light.py time 1000 5000 2>&1 | grep alignment_scores
  ALAD_CTA_GROUP=$cg timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('bench5k ms/step', round(d['ms_per_step'],2), 'kernel ms', round(r['avg_launch_ms'],2), 'TF', round(r['achieved'],1), 'clk', d['clocks']['sm_mhz'], 'R@1', d['recall_at_1'])
"
done
