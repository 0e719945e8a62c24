import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pde_solver_b200 as P
from pde_solver_b200 import _lib
ctx = _lib.default_context()
def run(tag, n, faces, maxlev=None, kind="heat"):
    if maxlev: os.environ["PDE_B200_MAX_LEVELS"] = str(maxlev)
    else: os.environ.pop("PDE_B200_MAX_LEVELS", None)
    L = [1.0, 0.5, 0.25]
    p = _lib.op_params(kind, 3, n, L, 1.0, 0.1, 1.2e11, 8e10, bc=_lib.make_bc(faces))
    nv, _ = _lib.mesh_counts(3, n)
    nc = 3 if kind == "elasticity" else 1
    rng = np.random.default_rng(1)
    b = rng.standard_normal((nc, nv))
    x, st = _lib.op_solve(ctx, p, b, _lib.make_opts(rtol=1e-10, precond="gmg", max_iters=200))
    print(f"{tag:40s} n={n} maxlev={maxlev} levels={st['levels']} iters={st['iters_total']} conv={st['converged']} true={st['true_relres']:.2e}", flush=True)
LR = {0: 0.0, 1: 0.0}
ALL = {f: 0.0 for f in range(6)}
run("nat 16,8,4", [16,8,4], LR)
run("nat 16,8,4 maxlev2", [16,8,4], LR, 2)
run("nat 12,12,12 (3 lev, no n=1)", [12,12,12], LR)
run("nat 24,24,24 (4 lev)", [24,24,24], LR)
run("nat 16,8,8 (ends 2,1,1)", [16,8,8], LR)
run("nat 16,8,8 maxlev3 (ends 4,2,2)", [16,8,8], LR, 3)
run("nat 8,8,2 maxlev2 (ends 4,4,1)", [8,8,2], LR, 2)
run("dir 8,8,8", [8,8,8], ALL)
run("x0 only 16,16,16", [16,16,16], {0: 0.0})
run("x0 only 16,16,16 maxlev2", [16,16,16], {0: 0.0}, 2)
run("x0 only 16,16,16 maxlev3", [16,16,16], {0: 0.0}, 3)
run("elast x0 16,16,16", [16,16,16], {0: 0.0}, None, "elasticity")
run("elast x0 16,16,16 maxlev2", [16,16,16], {0: 0.0}, 2, "elasticity")
run("elast x0 12,12,12", [12,12,12], {0: 0.0}, None, "elasticity")
run("elast all 16,16,16", [16,16,16], ALL, None, "elasticity")
