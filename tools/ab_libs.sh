#!/bin/bash
# Same-box A/B of two builds of the C-ABI library on the default bench step (device-resident part only):
# aladin_b200/libalad_b200_prev.so (another commit, tools/build_prev.sh) against aladin_b200/libalad_b200.so, alternating.
# Usage on the box: bash tools/ab_libs.sh [passes]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pci.bus_id --format=csv,noheader | tee gpurun_out/ab_box.txt
for i in $(seq 1 ${1:-2}); do
  for lib in libalad_b200_prev.so libalad_b200.so; do
    ALAD_B200_LIB=$PWD/aladin_b200/$lib timeout 200 python bench.py --no-e2e --no-also --no-cpu-baseline --no-cublas-probe --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib pass $i: step %.2f ms  kernel %.2f ms  %.1f TFLOP/s  sm %s MHz' % (d['ms_per_step'], r['avg_launch_ms'], r['achieved'], d['clocks']['sm_mhz']))"
  done
done | tee gpurun_out/ab_libs.log
