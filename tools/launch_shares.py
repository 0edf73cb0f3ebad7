#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    name = re.sub(r"\s+", " ", r["Kernel Name"])[:150]
    tot[name][0] += us
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
n = sum(v[1] for v in tot.values())
print(f"# launches {n} total {total/1e3:.1f} ms")
for name, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{us/1e3:9.2f} ms {100*us/total:5.1f}% x{c:4d} avg {us/c:8.1f} us  {name}")
