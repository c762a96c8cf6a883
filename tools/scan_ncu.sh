#!/bin/bash
# ncu evidence for the scan-sentences path: launch list (time + DRAM bytes) of one forward + backward at B = 512, and one
# --set full capture of the forward pair kernel.  Usage on the box: bash tools/scan_ncu.sh
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/scan_launches.csv python tools/scan_probe.py one 512 > gpurun_out/scan_ncu.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_pair_kernel -c 1 -o gpurun_out/scan_pair_fwd \
    python tools/scan_probe.py one 512 >> gpurun_out/scan_ncu.log 2>&1; echo "full capture exit $?"
ncu -i gpurun_out/scan_pair_fwd.ncu-rep --page details 2>/dev/null | grep -E "Duration|Registers|Theoretical Occupancy|Achieved Occupancy|DRAM Throughput|Memory Throughput|Compute \(SM\)|L1/TEX Hit|L2 Hit|Warp Cycles Per Issued|Issued Ipc|No Eligible|Shared Memory Configuration|Dynamic Shared|Bank conflicts|Local" | head -40
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/scan_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
idi = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[idi], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (i, k), m in per.items():
    a = agg.setdefault(k.split("(")[0][:70], [0, 0.0, 0.0])
    a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]/1e3:10.1f} us {100*a[1]/tot:5.1f}%  n={a[0]:3d}  dram={a[2]/1e6:9.1f} MB  {k}")
PY
