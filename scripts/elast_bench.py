"""Time the full elasticity solve (GMG-PCG + von Mises projection) through the host API.

    python scripts/elast_bench.py 320 64 64 [--precond gmg] [--rtol 1e-10]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pde_solver_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int, nargs=3)
ap.add_argument("--precond", default="gmg")
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
nx, ny, nz = args.n
for rep in range(args.reps):
    t0 = time.perf_counter()
    f = P._solve_elasticity_3d_static(1.0, 0.2, 0.2, nx, ny, nz, 210e9, 0.3, 0.0, 0.0, -76518.0, "stress",
                                      rtol=args.rtol, precond=args.precond, as_arrays=True)
    wall = time.perf_counter() - t0
    st = P.last_stats()
    out = {"n": args.n, "ndofs": st["ndofs"], "iters": st["iters_total"], "levels": st["levels"],
           "converged": st["converged"], "relres": st["final_relres"], "solve_ms": st["solve_ms"],
           "proj_ms": st["projection"]["solve_ms"], "proj_iters": st["projection"]["iters_total"],
           "launches": st["launches"], "wall_s": wall, "vm_max": float(f.values.max()),
           "gdof_iters_per_s": st["ndofs"] * st["iters_total"] / st["solve_ms"] / 1e6,
           "env": {k: v for k, v in os.environ.items() if k.startswith("PDE_B200_")}}
    print(json.dumps(out), flush=True)
