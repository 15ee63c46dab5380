#!/usr/bin/env python
"""ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`) -> markdown table
kernel | launches | total ms | share.  Usage: python tools/launch_list_summary.py X.csv OUT.md "title line" """
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    name = re.sub(r"^(void )?(og::)?", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else (v if unit in ("ms", "msecond") else v * 1e3))
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
with open(sys.argv[2], "w") as f:
    f.write(sys.argv[3] + "\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / total:.1f} % |\n")
print(f"{len(tot)} kernels, {total:.1f} ms")
