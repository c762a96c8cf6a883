#!/bin/bash
# GPU-box pass after the ranking / ListNet sweep merges: GPU parity tests, the HBM probe (CUDA-event times +
# equality checks), the default bench line.  Usage: bash tools/hbm_round.sh   -- everything lands in gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r02b_pytest.log
timeout 200 python tools/hbm_probe.py gpurun_out/r02b_hbm_probe.json > gpurun_out/r02b_hbm_probe.log 2>&1; echo "probe exit $?"
tail -70 gpurun_out/r02b_hbm_probe.log
timeout 300 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench exit $?"
cat gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
