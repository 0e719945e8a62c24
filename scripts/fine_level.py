#!/usr/bin/env python3
"""Per-kernel durations of the FINE level from an ncu launch list (gpu__time_duration.sum csv): launches within 50 % of the
longest launch of a kernel count as fine-level launches.  usage: scripts/fine_level.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    d[name].append(v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    big = sorted(v, reverse=True)
    fine = [x for x in big if x > 0.5 * big[0]]
    print(f"{k:34s} n={len(v):5d} share={100 * sum(v) / tot:5.1f}%  fine n={len(fine):4d} avg={sum(fine) / len(fine):8.1f} us"
          f"  coarser total={sum(v) - sum(fine):9.1f} us")
