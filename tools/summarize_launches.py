"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], val * scale))
agg = defaultdict(lambda: [0.0, 0])
for name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    agg[short][0] += us
    agg[short][1] += 1
total = sum(v[0] for v in agg.values())
print(f"# launch summary of {path}: {len(rows)} launches, {total/1e3:.3f} ms total (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total us | share |")
print("|---|---:|---:|---:|")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{name[:110]}` | {n} | {us:.1f} | {100*us/total:.1f}% |")
