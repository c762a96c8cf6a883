#!/bin/bash
# One GPU-box pass: GPU parity tests, the default bench line, ncu launch lists (bench step + loss kernels), training-step numbers.
# Usage (from the repo root on the box): bash tools/gpu_round.sh   -- everything lands in gpurun_out/
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cublas-probe > gpurun_out/bench_ncu.log 2>&1; echo "ncu bench exit $?"
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/loss_launches.csv \
    python tools/loss_probe.py > gpurun_out/loss_ncu.log 2>&1; echo "ncu loss exit $?"
timeout 300 python tools/bench_train_step.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "train step exit $?"
cat gpurun_out/train_step.json
