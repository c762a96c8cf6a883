#!/bin/bash
# scan-sentences: GPU parity tests, then timings with the register-resident forward kernel and with the shared-memory one.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_scan_sentences.py -m gpu -q > gpurun_out/pytest_scan.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_scan.log
ALAD_SCAN_SMEM_FWD=1 timeout 300 python -m pytest tests/test_gpu_zz_scan_sentences.py -m gpu -q 2>&1 | tail -1
timeout 200 python tools/scan_probe.py > gpurun_out/scan_probe_reg.json 2>gpurun_out/scan_probe.err; echo "probe exit $?"
ALAD_SCAN_SMEM_FWD=1 timeout 200 python tools/scan_probe.py > gpurun_out/scan_probe_smem.json 2>>gpurun_out/scan_probe.err
python - <<'PY'
import json
for tag in ("reg", "smem"):
    d = json.load(open(f"gpurun_out/scan_probe_{tag}.json"))["scan_probe"]
    for r in d:
        if r["aggregation"] == "scan-sentences":
            print(tag, r["B"], r["precision"], "fwd %.3f ms  fwd+bwd %.3f ms" % (r["fwd_ms"], r["fwd_bwd_ms"]))
PY
