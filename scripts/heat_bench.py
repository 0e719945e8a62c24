"""Time the device-resident heat stepper: python scripts/heat_bench.py DIM NX [NY [NZ]] --steps K"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_solver_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("dim", type=int)
ap.add_argument("n", type=int, nargs="+")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--precond", default="auto")
args = ap.parse_args()
ctx = _lib.default_context()
bc = _lib.make_bc({f: 0.0 for f in range(2 * args.dim)})
hs = _lib.HeatStepper(ctx, args.dim, args.n, [1.0] * args.dim, 1.0, 0.01, T_initial=20.0, bc=bc,
                      opts=_lib.make_opts(precond=args.precond))
hs.step(2)
st = hs.step(args.steps)
nd = 1
for k in args.n:
    nd *= k + 1
print(json.dumps({"dim": args.dim, "n": args.n, "ndofs": nd, "ms_per_step": st["solve_ms"] / args.steps,
                  "iters_per_step": st["iters_total"] / args.steps, "levels": st["levels"],
                  "converged": st["converged"], "gdofs": nd * args.steps / st["solve_ms"] / 1e6,
                  "launches_per_step": st["launches"] / args.steps}))
hs.close()
