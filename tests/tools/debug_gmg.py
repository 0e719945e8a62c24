import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pde_solver_b200 as P
from oracle import fem_oracle as fo

def heat(tag, dim, L, n, okw, args, kw, precond):
    ref = fo.solve_heat(dim, L, n, **okw)
    fn = {1: P._solve_heat_1d_raw, 2: P._solve_heat_2d_raw, 3: P._solve_heat_3d_raw}[dim]
    f = fn(*args, **kw, precond=precond, as_arrays=True)
    st = P.last_stats()
    print(tag, precond, "err", fo.rel_l2(f.values[-1], ref.values[-1]), "iters", st["iters_total"], "levels", st["levels"],
          "relres", st["final_relres"], "conv", st["converged"], flush=True)

for pc in ("jacobi", "gmg"):
    heat("heat16 dirichlet", 3, [1,1,1], [16,16,16], dict(diffusivity=1.0, T_initial=20.0, dt=0.01, num_steps=5),
         (1,1,1,16,16,16,1.0,0.0,20.0,0.01,5), {}, pc)
    heat("heat LR natural", 3, [1,0.5,0.25], [16,8,4], dict(diffusivity=2.0, T_initial=1.0, dt=0.05, num_steps=1, T_left=10.0, T_right=1.0),
         (1,0.5,0.25,16,8,4,2.0,0.0,1.0,0.05,1), dict(T_left=10.0, T_right=1.0), pc)
    heat("heat LR natural 2lev", 3, [1,0.5,0.25], [6,6,6], dict(diffusivity=2.0, T_initial=1.0, dt=0.05, num_steps=1, T_left=10.0, T_right=1.0),
         (1,0.5,0.25,6,6,6,2.0,0.0,1.0,0.05,1), dict(T_left=10.0, T_right=1.0), pc)
    ref = fo.solve_elasticity(3, [1, 0.2, 0.2], [20, 4, 4], 210e9, 0.3, body=[0, 0, -76518.0])
    g = P._solve_elasticity_3d_static(1, 0.2, 0.2, 20, 4, 4, 210e9, 0.3, 0.0, 0.0, -76518.0, "stress", precond=pc, as_arrays=True)
    st = P.last_stats()
    print("elast 20x4x4", pc, "err", fo.rel_l2(g.values[0], ref.values[0]), "iters", st["iters_total"], "levels", st["levels"],
          "relres", st["final_relres"], "proj", st["projection"]["iters_total"], st["projection"]["final_relres"], flush=True)
