#!/usr/bin/env python3
"""profiles/r02_kernel_table.json from ncu launch lists (gpu__time_duration.sum): per kernel its share of the launch
list, the fine-level time per launch and the fraction of the measured HBM peak at the kernel's algorithmic bytes per dof.

    scripts/kernel_table.py heat=profiles/r02_launches_heat_cfg4.csv:135005697 elast=profiles/r02_launches_elast_cfg5.csv:253826307
(name=csv:dofs of the fine level).  bench.py embeds the file as `kernel_table`."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.0
# algorithmic bytes per dof of the fine-level launch
BYTES = [
    (r"k_sweep3d<1, 4, 0", 16, "heat operator apply (+ fused p.Ap)"),
    (r"k_sweep3d<1, 4, 1", 24, "heat residual"),
    (r"k_sweep3d<1, 4, 2", 24, "heat Chebyshev sweep (restart form; 32 with x_prev)"),
    (r"k_sweep3d<1, 4, 3", 16, "heat first two sweeps fused (zero guess)"),
    (r"k_heat_post2<64, 4, 0", 24, "heat two post-smoothing sweeps in one pass"),
    (r"k_heat_post2<64, 4, 1", 17, "heat residual + restriction in one pass"),
    (r"k_post2<", 24, "round-1 fused post sweeps"),
    (r"k_elast3d<0", 16, "elasticity operator apply (+ fused p.Ap)"),
    (r"k_elast3d<3", 16, "elasticity first two sweeps fused (zero guess)"),
    (r"k_elast3d<1", 24, "elasticity residual"),
    (r"k_elast3d<2, 0", 24, "elasticity Chebyshev sweep (restart / zero x_prev)"),
    (r"k_elast3d<2, 1", 32, "elasticity Chebyshev sweep with x_prev"),
    (r"k_cg_pupdate_flat", 40, "p = z + beta p with the deferred x += alpha p"),
    (r"k_cg_update_flat", 24, "r -= alpha q, r.r"),
    (r"k_prolong_add<", 17, "prolongation + correction"),
    (r"k_restrict<", 9, "restriction of the residual"),
    (r"k_cheby_first<", 16, "first smoother sweep from a zero guess (x = s D^-1 b)"),
    (r"k_face_rows<", None, "rows on natural faces (surface only)"),
    (r"k_cell_rhs<", 32, "von Mises load vector of project()"),
    (r"k_cg_update<", 48, "Jacobi-PCG update (projection solve)"),
    (r"k_cg_pupdate<", 24, "Jacobi-PCG p update (projection solve)"),
]


def table(path, dofs):
    rows = list(csv.reader(open(path)))
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
        d[re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")].append(v)
    tot = sum(sum(v) for v in d.values())
    out = []
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / tot < 0.004:
            continue
        big = sorted(v, reverse=True)
        fine = [x for x in big if x > 0.5 * big[0]]
        us = sum(fine) / len(fine)
        e = {"kernel": k, "share": round(sum(v) / tot, 4), "launches": len(v), "fine_level_us": round(us, 1)}
        for pat, b, what in BYTES:
            if k.startswith(pat.replace("\\", "")) or re.match(pat, k):
                e["what"] = what
                if b:
                    # scalar kernels of the vector solve (projection) run on dofs / 3 nodes
                    scalar = k.endswith("<1>") or k.startswith("k_sweep3d<1") or k.startswith("k_face_rows<1")
                    nd = dofs / 3 if (scalar and dofs % 3 == 0 and "elast" in path) else dofs
                    if "k_cell_rhs" in k:
                        nd = dofs / 3
                    e["bytes_per_dof"] = b
                    e["gbs"] = round(b * nd / us / 1e3, 0)
                    e["frac_of_measured_peak"] = round(b * nd / us / 1e3 / PEAK, 3)
                break
        out.append(e)
    return {"source": os.path.relpath(path, ROOT), "fine_level_dofs": dofs, "total_ms": round(tot / 1e3, 2), "kernels": out}


res = {"_comment": "shares of the ncu launch list (cold-cache, serialised: shares, not absolutes); frac = algorithmic bytes / "
                   "fine-level time / MEASURED_PEAKS hbm_gbs", "peak_gbs": PEAK}
for arg in sys.argv[1:]:
    name, rest = arg.split("=")
    path, dofs = rest.split(":")
    res[name] = table(os.path.join(ROOT, path) if not os.path.isabs(path) else path, int(dofs))
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_kernel_table.json"), "w"), indent=1)
print(json.dumps(res, indent=1)[:3000])
