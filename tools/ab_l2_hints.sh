#!/bin/bash
# A/B of the TMA L2 eviction hints (ALAD_L2_HINTS=0|1) x region-block size (ALAD_L2_BLOCK_MB):
# kernel time from bench.py (CUDA events) and DRAM bytes of one launch from a 1-pass ncu run.
summ='
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("  bench: ms/step", round(d["ms_per_step"],2), "kernel ms", round(r["avg_launch_ms"],2), "TF", round(r["achieved"],1), "clk", d["clocks"]["sm_mhz"])
'
for cfg in "1 30" "0 30" "1 60" "1 15" "1 100" "1 30" "0 30"; do
  set -- $cfg
  echo "=============== ALAD_L2_HINTS=$1 ALAD_L2_BLOCK_MB=$2"
  ALAD_L2_HINTS=$1 ALAD_L2_BLOCK_MB=$2 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 6 2>/dev/null | python -c "$summ"
done
for cfg in "1 30" "0 30" "1 60" "1 100"; do
  set -- $cfg
  echo "=============== ncu dram bytes: ALAD_L2_HINTS=$1 ALAD_L2_BLOCK_MB=$2"
  ALAD_L2_HINTS=$1 ALAD_L2_BLOCK_MB=$2 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:mrsw_fwd -c 1 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 0 2>&1 | grep -E "dram__|gpu__time|lts__|sm__pipe"
done
