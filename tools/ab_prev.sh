#!/bin/bash
# Same-box A/B of two builds of the library (aladin_b200/libalad_b200_prev.so = another commit, built by tools/build_prev.sh):
# prefetch on / off at block 74, three rounds each, alternating the builds.  Usage on the box: bash tools/ab_prev.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pci.bus_id,serial --format=csv | tee gpurun_out/ab_box.txt
for i in 1 2; do
  for lib in libalad_b200_prev.so libalad_b200.so; do
    echo "== $lib (pass $i)"
    ALAD_B200_LIB=$PWD/aladin_b200/$lib SWEEP_CONFIGS="${AB_CONFIGS:-1074:4,1074:0}" timeout 200 python tools/sweep_tile_order.py 3 2>&1 | tail -1
  done
done | tee gpurun_out/ab_prev.log
