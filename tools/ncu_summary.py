"""Turn an `ncu --csv --log-file` launch list (metrics gpu__time_duration.sum and, optionally,
dram__bytes_read.sum / dram__bytes_write.sum) into the summaries kept under profiles/:
    python tools/ncu_summary.py launches.csv --shares  > profiles/rNN_launches_*.csv     (kernel, launches, total ms, share)
    python tools/ncu_summary.py launches.csv --hbm [--only alad] [--peak 6544.3]         (markdown table, one row per launch)"""
import argparse
import csv
import io
import re
import sys
from collections import OrderedDict

UNIT = {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "second": 1.0, "s": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(path):
    """-> list of dicts {name, time_s, rd, wr} in launch order."""
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit(f"{path}: no ncu CSV header found")
    rows = OrderedDict()
    for r in csv.DictReader(io.StringIO(text[start:])):
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except (ValueError, KeyError, TypeError):
            continue
        e = rows.setdefault(r["ID"], {"name": r["Kernel Name"], "time_s": 0.0, "rd": None, "wr": None})
        scale = UNIT.get(r.get("Metric Unit", ""), 1.0)
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            e["time_s"] = val * scale
        elif m.startswith("dram__bytes_read"):
            e["rd"] = val * scale
        elif m.startswith("dram__bytes_write"):
            e["wr"] = val * scale
    return list(rows.values())


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--shares", action="store_true")
    ap.add_argument("--hbm", action="store_true")
    ap.add_argument("--only", default=None, help="substring filter on kernel names")
    ap.add_argument("--peak", type=float, default=6544.3, help="HBM copy peak, GB/s (MEASURED_PEAKS.json)")
    a = ap.parse_args()
    L = launches(a.csv)
    if a.only:
        L = [l for l in L if a.only in l["name"]]
    if a.shares:
        agg = OrderedDict()
        for l in L:
            e = agg.setdefault(short(l["name"]), [0, 0.0])
            e[0] += 1
            e[1] += l["time_s"]
        total = sum(e[1] for e in agg.values())
        print(f"# total device time of listed launches: {1e3 * total:.1f} ms")
        print("kernel,launches,total_ms,share")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{k},{n},{1e3 * t:.3f},{t / total:.4f}")
    if a.hbm:
        print("| kernel | duration | DRAM read | DRAM write | achieved GB/s | of measured HBM peak |")
        print("|---|---|---|---|---|---|")
        for l in L:
            if l["rd"] is None:
                continue
            gbs = (l["rd"] + l["wr"]) / l["time_s"] / 1e9 if l["time_s"] > 0 else 0.0
            print(f"| `{short(l['name'])}` | {1e6 * l['time_s']:.1f} us | {l['rd'] / 1e6:.1f} MB | {l['wr'] / 1e6:.1f} MB | "
                  f"{gbs:.0f} | {gbs / a.peak:.2f} |")


if __name__ == "__main__":
    sys.exit(main())
