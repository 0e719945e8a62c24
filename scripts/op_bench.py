"""Operator micro-benchmark: time device-resident applications of one matrix-free operator.

    python scripts/op_bench.py heat 512 512 512 [--reps 20]
Prints one JSON line (ms per apply, GDOF/s, algorithmic GB/s at 16 B/dof)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_solver_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("kind")
ap.add_argument("n", type=int, nargs=3)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--bc", default="all", choices=["all", "clamp", "none"])
args = ap.parse_args()
ctx = _lib.default_context()
lam, mu = 121.15e9, 80.77e9
faces = {"all": {f: 0.0 for f in range(6)}, "clamp": {0: 0.0}, "none": {}}[args.bc]
L = [1.0, 1.0, 1.0] if args.kind != "elasticity" else [1.0, 0.2, 0.2]
p = _lib.op_params(args.kind, 3, args.n, L, 1.0, 0.01, lam, mu, bc=_lib.make_bc(faces), variant=args.variant)
ms, nd = _lib.op_bench(ctx, p, reps=args.reps, warmup=3)
print(json.dumps({"kind": args.kind, "n": args.n, "bc": args.bc, "variant": args.variant, "ms": ms, "ndofs": nd,
                  "gdofs": nd / ms / 1e6, "GBps_16B": 16 * nd / ms / 1e6,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("PDE_B200_")}}))
