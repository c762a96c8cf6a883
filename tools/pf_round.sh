#!/bin/bash
# Is the word-row L2 prefetch a win on THIS box?  In-process rotation of (block 74 + prefetch), (74, no prefetch), (148, no prefetch),
# then ncu DRAM / L2 probes of the same three.  Usage on the box: bash tools/pf_round.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pci.bus_id,serial,vbios_version,clocks.max.sm,power.limit --format=csv | tee gpurun_out/pf_box.txt
SWEEP_CONFIGS="1074:4,1074:0,1148:0,1074:4,1074:0,1037:0" timeout 300 python tools/sweep_tile_order.py 3 > gpurun_out/sweep_pf.log 2>&1; echo "sweep exit $?"
tail -1 gpurun_out/sweep_pf.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second
for cfg in "74 1" "74 0" "148 0"; do
  set -- $cfg
  echo "=============== ALAD_N_BLOCK=$1 ALAD_L2_PREFETCH=$2"
  ALAD_N_BLOCK=$1 ALAD_L2_PREFETCH=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:mrsw_fwd -c 1 \
    python bench.py --no-cpu-baseline --no-e2e --no-cublas-probe --steps 1 --warmup 0 2>&1 | grep -E "dram__|gpu__time|lts__|sm__"
done > gpurun_out/ncu_pf_probe.log 2>&1
cat gpurun_out/ncu_pf_probe.log
