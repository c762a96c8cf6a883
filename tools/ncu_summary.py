"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + top stall sites.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread ",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum "]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for k, launch in enumerate(raw[2:]):
        print(f"== launch {k}: {launch[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
        for h, u, v in zip(hdr, units, launch):
            if any((h + ' ').startswith(key) or h == key.strip() for key in KEYS):
                print(f"  {h} [{u}] = {v}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    rows = src[2:]

    def g(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(g(r, "# Samples") for r in rows)
    print(f"== source: {len(rows)} SASS instructions, {int(tot)} stall samples; top {top}:")
    for r in sorted(rows, key=lambda r: -g(r, "# Samples"))[:top]:
        print(f"  {100 * g(r, '# Samples') / max(tot, 1):5.1f}%  exec={int(g(r, 'Instructions Executed')):>11d}  {r[ix['Source']][:90]}")


if __name__ == "__main__":
    main()
