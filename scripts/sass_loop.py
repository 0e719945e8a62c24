#!/usr/bin/env python3
"""Static instruction mix of the hot loop of a kernel, from `cuobjdump -sass` output (no GPU needed).

usage: scripts/sass_loop.py file.sass 'k_sweep3d<3, 2, 0, false>' [outputs_per_trip]

Finds every backward branch of the function, takes the loop with the largest span (the unrolled plane march) and
prints its opcode histogram, optionally divided by the number of outputs one trip of the loop produces."""
import collections
import re
import subprocess
import sys


def functions(path):
    cur, out = None, {}
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def main():
    path, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        loops = []
        for addr, text in ins:
            m = re.search(r"\bBRA(?:\.U)?\S*\s+(?:!?U?P\d,?\s+)*`\(\.L_x_\d+\)|BRA\S*\s.*?0x([0-9a-f]+)", text)
            m2 = re.search(r"BRA.*?(0x[0-9a-f]+)", text)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < addr:
                    loops.append((addr - tgt, tgt, addr))
        print(f"== {name[:100]}\n   {len(ins)} instructions, {len(loops)} backward branches")
        if not loops:
            continue
        # the plane march is unrolled three times with one CTA barrier per plane: prefer the tightest loop that holds
        # exactly `nbar` barriers (4th argument), otherwise list the widest loops
        nbar = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        if nbar:
            sel = [l for l in loops if sum(1 for a, t in ins if l[1] <= a <= l[2] and re.search(r"\bBAR\.SYNC", t)) == nbar]
            sel.sort(key=lambda l: -sum(1 for a, t in ins if l[1] <= a <= l[2] and "DFMA" in t))
            loops = sel[:1] if sel else sorted(loops, reverse=True)[:3]
        else:
            loops = sorted(loops, reverse=True)[:3]
        for span, lo, hi in loops:
            body = [t for a, t in ins if lo <= a <= hi]
            hist = collections.Counter()
            for t in body:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                op = t.split()[0]
                op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "LDG", "STG", "STS", "LDC")) else op.split(".")[0]
                hist[op] += 1
            n = len(body)
            print(f"   loop [{lo:#x}, {hi:#x}]: {n} instructions = {n / per:.1f} per output")
            for op, c in hist.most_common(28):
                print(f"      {op:14s} {c:5d}  {c / per:7.2f}/out")


if __name__ == "__main__":
    main()
