"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): per-kernel launches, total time and share.
Usage: python tools/summarize_launches.py launches.csv > summary.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    rows.append((name, us))
agg = OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"Total {tot / 1e3:.2f} ms over {len(rows)} launches.\n")
print("| kernel | launches | us total | share |\n|---|---:|---:|---:|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")
