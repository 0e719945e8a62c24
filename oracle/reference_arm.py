"""CPU reference arm for bench.py (test/bench infrastructure only).

Times what the reference's hot loop does on every backward-Euler step through DOLFIN/PETSc
(`solve(a == L_form, u_sol, bcs)`, fenics_mcp_server.py:709): assemble A, assemble b,
apply the Dirichlet rows, sparse direct LU factorisation + solve — restated with SciPy/SuperLU
because FEniCS is not installable here.  SciPy's assembly, SpMV and SuperLU are single-threaded."""
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem_oracle as fo


class HeatReference3D:
    def __init__(self, n, L=(1.0, 1.0, 1.0), kappa=1.0, dt=0.01, T_initial=20.0, T_boundary=0.0):
        self.n, self.kappa, self.dt = n, kappa, dt
        self.mesh = fo.make_mesh(3, list(L), [n, n, n])
        nv = self.mesh.nv
        nn = n + 1
        i = np.arange(nv)
        ix, iy, iz = i % nn, (i // nn) % nn, i // (nn * nn)
        self.bc = np.nonzero((ix == 0) | (ix == n) | (iy == 0) | (iy == n) | (iz == 0) | (iz == n))[0]
        self.g = np.full(self.bc.size, float(T_boundary))
        self.u = np.full(nv, float(T_initial))
        self.u[self.bc] = self.g
        self.ndofs = nv

    def step(self):
        """One reference time step: assemble, apply BCs row-wise, factorise, solve."""
        K, M = fo.assemble_stiffness_mass(self.mesh)                # assemble(a) pieces
        A = (M + (self.dt * self.kappa) * K).tocsr()
        b = M @ self.u                                              # assemble(L)
        A, b = fo.apply_bc_rowwise(A, b, self.bc, self.g)           # bc.apply(A, b)
        s = fo._row_scale(A)
        lu = spla.splu((sp.diags(s) @ A).tocsc(), permc_spec="MMD_AT_PLUS_A")   # LU every step
        self.u = lu.solve(s * b)
        return self.u

    def run(self, steps, warmup=0):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return time.perf_counter() - t0
