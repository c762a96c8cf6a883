#!/bin/bash
# Short GPU-box pass: all GPU parity tests (no -x: every failure is listed), smoke, 'scan-sentences' timings, the default bench line.
# Usage (from the repo root on the box): bash tools/gpu_verify.sh   -- everything lands in gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 240 python tools/scan_probe.py > gpurun_out/scan_probe.json 2> gpurun_out/scan_probe.err; echo "scan probe exit $?"
cat gpurun_out/scan_probe.json; tail -5 gpurun_out/scan_probe.err
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json
