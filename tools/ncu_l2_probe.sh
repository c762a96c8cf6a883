#!/bin/bash
# L2 / DRAM behaviour of one scoring launch for several region-block sizes (in tiles)
M=dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__d_sectors.avg.pct_of_peak_sustained_elapsed,lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed
for nb in "$@"; do
  echo "=============== ALAD_N_BLOCK=$nb"
  ALAD_N_BLOCK=$nb timeout 600 ncu --metrics $M --clock-control none -k regex:mrsw_fwd -c 1 \
    python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 0 2>&1 | grep -E "dram__|gpu__time|lts__|sm__"
done
