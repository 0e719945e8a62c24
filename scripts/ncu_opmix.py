"""Opcode mix (per output) of the largest kernel in an ncu report: python scripts/ncu_opmix.py rep.ncu-rep NOUT [regex]"""
import collections
import csv
import subprocess
import sys

rep, nout = sys.argv[1], float(sys.argv[2])
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 3:
    cmd += ["--kernel-name", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blocks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = [r[1]]
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
best = None
for b in blocks:
    hdr = b[1]
    ie = hdr.index("Instructions Executed")
    tot = sum(int(r[ie]) for r in b[2:] if len(r) > ie and r[ie].isdigit())
    if best is None or tot > best[0]:
        best = (tot, b)
tot, b = best
hdr = b[1]
ie, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
agg, st = collections.Counter(), collections.Counter()
for r in b[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    toks = r[isrc].split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
    agg[op] += int(r[ie])
    st[op] += int(r[ist])
print(b[0][:100])
print("warp instructions", tot, " thread instructions per output", round(tot * 32 / nout, 1))
S = max(1, sum(st.values()))
for op, n in agg.most_common(24):
    print(f"{op:14s} {n / tot * 100:6.2f}%  {n * 32 / nout:7.2f}/out  stall samples {st[op] / S * 100:5.1f}%")
