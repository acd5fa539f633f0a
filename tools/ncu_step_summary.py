"""Digest of an `ncu --page raw --csv` export of one training step: per kernel family the launch count, total
device time, DRAM bytes read+written per launch, DRAM throughput % and tensor-pipe utilisation %.

    ncu -i step.ncu-rep --page raw --csv > step_raw.csv
    python tools/ncu_step_summary.py step_raw.csv profiles/rNN_ncu_step_summary.md profiles/rNN_gemm_traffic.json
"""
import csv
import json
import re
import sys
from collections import defaultdict

src, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src, newline="")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    i = col.get(name)
    if i is None or i >= len(r):
        return None
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return None


def unit(name):
    i = col.get(name)
    return rows[1][i] if i is not None else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


agg = defaultdict(lambda: defaultdict(float))
for r in rows[2:]:
    if len(r) < len(hdr) // 2:
        continue
    name = r[col["Kernel Name"]]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)[:90]
    a = agg[name]
    a["n"] += 1
    a["us"] += to_us(val(r, "gpu__time_duration.sum") or 0, unit("gpu__time_duration.sum"))
    a["rd"] += to_bytes(val(r, "dram__bytes_read.sum") or 0, unit("dram__bytes_read.sum"))
    a["wr"] += to_bytes(val(r, "dram__bytes_write.sum") or 0, unit("dram__bytes_write.sum"))
    a["dram_pct"] += val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0
    a["tensor_pct"] += val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0
    a["regs"] = val(r, "launch__registers_per_thread") or 0

tot = sum(a["us"] for a in agg.values())
with open(out_md, "w") as f:
    f.write(f"# ncu --set full digest of one training step ({int(sum(a['n'] for a in agg.values()))} launches, "
            f"{tot / 1e3:.2f} ms serialised, cold caches: compare SHARES)\n\n")
    f.write("| kernel | launches | total us | share | DRAM MB/launch (rd+wr) | DRAM % of peak | tensor pipe % active | regs |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        f.write(f"| `{name}` | {int(n)} | {a['us']:.0f} | {100 * a['us'] / tot:.1f}% | {(a['rd'] + a['wr']) / n / 1e6:.1f} "
                f"| {a['dram_pct'] / n:.1f} | {a['tensor_pct'] / n:.1f} | {int(a['regs'])} |\n")

gemm = {k: v for k, v in agg.items() if "gemm_tcgen05_kernel" in k}
n = sum(v["n"] for v in gemm.values())
if n:
    js = {"kernel": "gemm_tcgen05_kernel (all instantiations of one step)", "launches": int(n),
          "dram_bytes_per_launch": (sum(v["rd"] + v["wr"] for v in gemm.values())) / n,
          "tensor_pipe_pct_active": sum(v["tensor_pct"] for v in gemm.values()) / n,
          "dram_pct_of_peak": sum(v["dram_pct"] for v in gemm.values()) / n,
          "source": src}
    json.dump(js, open(out_json, "w"), indent=1)
    print(js)
