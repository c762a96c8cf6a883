#!/bin/bash
# One ncu --set full capture of the scoring kernel inside the default bench step (the source of roofline.traffic in
# profiles/traffic.json).  Usage on the box: bash tools/ncu_full_fwd.sh
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mrsw_fwd -c 1 -o gpurun_out/mrsw_fwd_full -f \
    python bench.py --no-cpu-baseline --no-e2e --no-cublas-probe --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/mrsw_fwd_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/mrsw_fwd_full_raw.csv
ncu -i gpurun_out/mrsw_fwd_full.ncu-rep --page details 2>/dev/null > gpurun_out/mrsw_fwd_full_details.txt
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/mrsw_fwd_full_raw.csv')))
hdr, vals = rows[0], rows[-1]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
units = rows[1]
for w in want:
    if w in hdr:
        i = hdr.index(w); print(w, vals[i], units[i])
PY
