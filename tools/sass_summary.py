"""SASS opcode summary of the built library: per kernel, the counts of the Blackwell-specific instructions that prove the
tcgen05 / TMEM / TMA path (B200_PROFILING.md lists the mnemonics) plus registers and spills.
    python tools/sass_summary.py > profiles/rNN_sass_summary.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "aladin_b200", "libalad_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "HMMA",
         "FMNMX3", "FMNMX", "REDG", "RED", "ATOMG", "LDGSTS", "BAR", "MEMBAR", "SHFL", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    lines = res.splitlines()
    for i, l in enumerate(lines):
        m = re.search(r"Function (\S+):", l)
        if m and i + 1 < len(lines):
            u = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", lines[i + 1])
            if u:
                usage[m.group(1)] = tuple(map(int, u.groups()))
    kernels = OrderedDict()
    cur = None
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = kernels.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", l)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total"] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTMAPF"):
                cur["full:" + full] += 1
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode summary of `aladin_b200/libalad_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` (cuobjdump -sass / -res-usage of the in-tree build).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
    print("UTMALDG = TMA load (cp.async.bulk.tensor), UTCBAR = tcgen05.commit -> mbarrier, UTMAPF = TMA L2 prefetch, SYNCS = mbarrier ops.\n")
    print("| kernel | instr | regs | stack/local | " + " | ".join(WATCH[:11]) + " | FMNMX3 | RED/ATOM | SHFL |")
    print("|---|---|---|---|" + "---|" * 14)
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*$", "", name).replace("void ", "")
        u = usage.get(mangled, (0, 0, 0, 0))
        cells = [str(c.get(w, 0)) for w in WATCH[:11]]
        print(f"| `{short}` | {c['__total']} | {u[0]} | {u[1]}/{u[3]} | " + " | ".join(cells) +
              f" | {c.get('FMNMX3', 0)} | {c.get('REDG', 0) + c.get('RED', 0) + c.get('ATOMG', 0)} | {c.get('SHFL', 0)} |")
    print("\n## Variants of the tensor-core / TMA instructions in the scoring kernels\n")
    for (mangled, c), name in zip(kernels.items(), demangle):
        if "mrsw_fwd" not in name:
            continue
        short = re.sub(r"\(.*$", "", name).replace("void ", "")
        print(f"* `{short}`: " + ", ".join(f"{k[5:]} x{v}" for k, v in sorted(c.items()) if k.startswith("full:")))


if __name__ == "__main__":
    main()
