#!/bin/bash
# ncu --set full capture of ONE launch of the pair-list scoring kernel (two-stage retrieval, stage 2) at COCO-5k shape.
# Launch order inside one two_stage_retrieval call: stage-1 GEMM M, stage-1 GEMM Mt (both mrsw_fwd_kernel<2,false>,
# plain-GEMM epilogue), then the pair-list launch (mrsw_fwd_kernel<1,true>) -> skip 2 launches of the name.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mrsw_fwd_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/pairs_full -f python tools/two_stage_probe.py > gpurun_out/ncu_pairs.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/pairs_full.ncu-rep --page details > gpurun_out/pairs_full_details.txt 2>&1
ncu -i gpurun_out/pairs_full.ncu-rep --page raw --csv > gpurun_out/pairs_full_raw.csv 2>&1
