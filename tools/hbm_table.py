"""Markdown table of per-launch duration and DRAM bytes from an ncu CSV made with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X ...
usage: python tools/hbm_table.py X.csv [peak_GBs] [name-filter-regex]"""
import csv
import re
import sys


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * mult.get(unit, 1)


def to_us(v, unit):
    mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    return float(v.replace(",", "")) * mult.get(unit, 1)


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6556.2
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        key = int(r["ID"])
        d = launches.setdefault(key, {"name": r["Kernel Name"]})
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_read.sum":
            d["rd"] = to_bytes(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_bytes(r["Metric Value"], r["Metric Unit"])
    print("| id | kernel | duration | DRAM read | DRAM write | achieved GB/s | of measured HBM peak |")
    print("|---|---|---|---|---|---|---|")
    for key in sorted(launches):
        d = launches[key]
        name = re.sub(r"\(.*", "", d["name"]).replace("alad::", "").strip()
        if pat and not pat.search(name):
            continue
        if "us" not in d:
            continue
        rd, wr = d.get("rd", 0.0), d.get("wr", 0.0)
        gbs = (rd + wr) / (d["us"] * 1e-6) / 1e9
        print(f"| {key} | `{name}` | {d['us']:.1f} us | {rd / 1e6:.1f} MB | {wr / 1e6:.1f} MB | {gbs:.0f} | {gbs / peak:.2f} |")


if __name__ == "__main__":
    main()
