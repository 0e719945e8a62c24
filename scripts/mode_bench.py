"""Time every sweep mode of one operator (device-resident): apply, residual, Chebyshev restart / with x_prev / zero x_prev.

    python scripts/mode_bench.py elasticity 1280 256 256 [--bc clamp] [--reps 10]
One JSON line per mode: ms per launch and algorithmic GB/s (16 / 24 / 24 / 32 / 24 B per dof)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_solver_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("kind")
ap.add_argument("n", type=int, nargs=3)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--bc", default=None, choices=["all", "clamp", "none"])
ap.add_argument("--modes", default="0,1,2,3,4")
args = ap.parse_args()
ctx = _lib.default_context()
lam, mu = 121.15e9, 80.77e9
bcname = args.bc or ("clamp" if args.kind == "elasticity" else "all")
faces = {"all": {f: 0.0 for f in range(6)}, "clamp": {0: 0.0}, "none": {}}[bcname]
L = [1.0, 1.0, 1.0] if args.kind != "elasticity" else [1.0, 0.2, 0.2]
p = _lib.op_params(args.kind, 3, args.n, L, 1.0, 0.01, lam, mu, bc=_lib.make_bc(faces))
BYTES = {0: 16, 1: 24, 2: 24, 3: 32, 4: 24}
NAME = {0: "apply", 1: "residual", 2: "cheby_restart", 3: "cheby_prev", 4: "cheby_zero_prev"}
for mode in [int(m) for m in args.modes.split(",")]:
    ms, nd = _lib.op_bench_mode(ctx, p, mode, reps=args.reps, warmup=3)
    print(json.dumps({"kind": args.kind, "n": args.n, "bc": bcname, "mode": NAME[mode], "ms": round(ms, 4), "ndofs": nd,
                      "B_per_dof": BYTES[mode], "GBps": round(BYTES[mode] * nd / ms / 1e6, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("PDE_B200_")}}), flush=True)
