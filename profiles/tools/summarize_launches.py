"""Fold an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: python profiles/tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/rNN_launch_summary.csv"""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mi = hdr.index("Metric Name") if "Metric Name" in hdr else -1      # multi-metric logs: only the duration rows count
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= vi or not r[vi] or (mi >= 0 and r[mi] != "gpu__time_duration.sum"):
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    t = tot[name]
    t[0] += 1
    t[1] += ns
total = sum(t[1] for t in tot.values())
n = sum(t[0] for t in tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print(f"# total device time {total * 1e-6:.1f} ms over {n} launches; per-launch times are cold-cache/serialised: compare shares")
print("kernel,launches,total_ms,share_pct,avg_us")
for name, (cnt, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{cnt},{ns * 1e-6:.3f},{100 * ns / total:.2f},{ns / cnt * 1e-3:.2f}")
