#!/bin/bash
# First GPU pass of the next round (1 GPU, ~4 GPU-minutes): what the last day of round 1 changed but could not re-measure.
#   1. parity + default bench line + ncu launch list of the 4-stage / no-prefetch build (profiles/ still hold the 6-stage list)
#   2. per-rank shapes of the 2/4/8-GPU runs (2500 / 1250 / 625 images x 25000 captions) with the new ring: block sizes 74 and
#      the alternatives that won with the 6-stage ring (profiles/r01_tile_order_sweep.md sections 5, 6)
#   3. e2e from host tensors: phase count of the upload (the device-resident step is now 10 ms faster than e2e)
# Usage on the box: bash tools/next_round_gpu.sh      (everything lands in gpurun_out/)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1.json
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cublas-probe > gpurun_out/bench_ncu.log 2>&1; echo "ncu launch list exit $?"
for ni in 2500 1250 625; do
  echo "== per-rank shape: $ni images"
  SWEEP_CONFIGS="1074:0,1037:0,1045:0,1090:0" timeout 200 python tools/sweep_tile_order.py 3 $ni 25000 2>&1 | tail -1
done | tee gpurun_out/sweep_per_rank.log
