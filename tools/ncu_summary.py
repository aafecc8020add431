#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if u in ("ns", "nsecond"):
        v /= 1000
    elif u in ("ms", "msecond"):
        v *= 1000
    elif u in ("s", "second"):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%-58s %7s %12s %10s %6s" % ("kernel", "launches", "total_us", "avg_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %7d %12.1f %10.2f %5.1f%%" % (k[:58], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print("%-58s %7d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
