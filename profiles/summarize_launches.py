#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 2:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us of kernel time (ncu: cold cache, serialised)")
print(f"{'kernel':34s} {'launches':>8s} {'total us':>10s} {'mean us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / v[0] / 1e3:9.2f} {100 * v[1] / tot:6.1f}%")
