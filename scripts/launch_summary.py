"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (name, grid size)."""
import collections
import csv
import io
import sys

path = sys.argv[1]
bygrid = len(sys.argv) > 2
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    name = row["Kernel Name"].split("(")[0][:50]
    key = (name, row.get("Grid Size", "")) if bygrid else (name,)
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{' '.join(k):70s} n={n:4d} total={t/1e3:9.3f} ms avg={t/n:9.1f} us {100*t/tot:5.1f}%")
print(f"total {tot/1e3:.3f} ms")
