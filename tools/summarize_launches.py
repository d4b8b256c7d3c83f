"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table
(dev tool).  usage: python tools/summarize_launches.py launches.csv [live_ms_per_step] > summary.md"""
import csv, re, sys
from collections import defaultdict

rows = []
hdr = None
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if hdr is None:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:110]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
# ncu prints nested namespaces without the outer one: unimp::lm::x -> lm::x, unimp::ff::x -> ff::x
ours = {k: v for k, v in agg.items() if "unimp::" in k or re.search(r"\b(lm|ff)::", k)}
print(f"launches: {n}; sum of durations {tot / 1e3:.2f} ms; our kernels: {sum(v[0] for v in ours.values())} launches, "
      f"{sum(v[1] for v in ours.values()) / 1e3:.2f} ms ({100 * sum(v[1] for v in ours.values()) / tot:.1f} % of the sum)"
      + (f"; bench.py's live number for the same step: {sys.argv[2]} ms" if len(sys.argv) > 2 else ""))
print()
print("| kernel | launches | total ms | avg us | share |")
print("|---|---:|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"| `{k}` | {c} | {us / 1e3:.3f} | {us / c:.2f} | {100 * us / tot:.1f} % |")
