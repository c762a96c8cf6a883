#!/bin/bash
# Same-box A/B of the ranking sweep with 8 (default build) and 16 (libalad_b200_vec4.so, -DALAD_FS_VEC=4) columns per thread:
# ranking / retrieval / loss parity tests on both builds, the HBM probe on both, ncu launch lists (durations + DRAM bytes).
mkdir -p gpurun_out
nvidia-smi -L
V4=$PWD/aladin_b200/libalad_b200_vec4.so
timeout 200 python -m pytest tests/test_gpu_ranking.py tests/test_gpu_retrieval.py tests/test_gpu_losses.py -m gpu -q 2>&1 | tail -3
ALAD_B200_LIB=$V4 timeout 200 python -m pytest tests/test_gpu_ranking.py tests/test_gpu_retrieval.py -m gpu -q 2>&1 | tail -3
timeout 100 python tools/hbm_probe.py gpurun_out/r02c_hbm_probe.json > gpurun_out/r02c_hbm_probe.log 2>&1; echo "probe exit $?"
ALAD_B200_LIB=$V4 timeout 100 python tools/hbm_probe.py gpurun_out/r02c_hbm_probe_vec4.json > gpurun_out/r02c_hbm_probe_vec4.log 2>&1; echo "probe vec4 exit $?"
python - <<'PY'
import json
for n in ("r02c_hbm_probe", "r02c_hbm_probe_vec4"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        r = d["ranking"]
        print(n, {k: round(v, 4) if isinstance(v, float) else v for k, v in r.items() if k.endswith("_ms") or k.endswith("equal") or k.endswith("separate")})
        print("   listnet", [(l["B"], round(l["listnet_ms"], 4), round(l["triplet_ms"], 4)) for l in d["losses"]])
    except Exception as e:
        print(n, "unreadable", e)
PY
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 120 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02c_hbm_launches.csv python tools/hbm_ncu.py > /dev/null 2>&1; echo "ncu exit $?"
ALAD_B200_LIB=$V4 timeout 120 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02c_hbm_launches_vec4.csv python tools/hbm_ncu.py > /dev/null 2>&1; echo "ncu vec4 exit $?"
grep -h "rank_sweep_kernel" gpurun_out/r02c_hbm_launches.csv gpurun_out/r02c_hbm_launches_vec4.csv | grep duration | cut -c1-40,200-
