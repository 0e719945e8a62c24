"""CPU reference arm for bench.py (test/bench infrastructure only).

Times what the reference's hot loop does on every backward-Euler step through DOLFIN/PETSc
(`solve(a == L_form, u_sol, bcs)`, fenics_mcp_server.py:709): assemble A, assemble b,
apply the Dirichlet rows, sparse direct LU factorisation + solve — restated with SciPy/SuperLU
because FEniCS is not installable here.  SciPy's assembly, SpMV and SuperLU are serial (BLAS calls inside may thread);
bench.py reports the measured CPU time / wall time as `cores`."""
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


class HeatCsrCg3D(HeatReference3D):
    """The same time step with a Krylov solver instead of LU: A = M + dt*kappa*K assembled
    ONCE into CSR (the set-up is not timed), symmetric Dirichlet elimination, Jacobi-PCG to rtol 1e-10 from the warm
    start u_n - what `solve(..., solver_parameters={"linear_solver": "cg", "preconditioner": "jacobi"})` would do
    per step.  SciPy's CSR SpMV is serial, NumPy's vector operations may use several BLAS threads: bench.py reports
    the measured CPU time / wall time as `cores`."""

    def __init__(self, n, rtol=1e-10, **kw):
        super().__init__(n, **kw)
        self.rtol = rtol
        K, M = fo.assemble_stiffness_mass(self.mesh)
        self.M = M.tocsr()
        A = (M + (self.dt * self.kappa) * K).tocsr()
        keep = np.ones(self.ndofs)
        keep[self.bc] = 0.0
        self.keep = keep
        self.A = A                       # rows / columns of the Dirichlet nodes are handled through `keep`
        self.dinv = keep / A.diagonal()
        self.iters = 0

    def step(self):
        b = self.M @ self.u
        x = self.u.copy()                # warm start; carries the Dirichlet values
        r = self.keep * (b - self.A @ x)
        bn = np.sqrt(np.dot(self.keep * b, self.keep * b))
        z = self.dinv * r
        p = z.copy()
        rz = np.dot(r, z)
        it = 0
        while np.sqrt(np.dot(r, r)) > self.rtol * bn and it < 100000:
            q = self.keep * (self.A @ p)
            alpha = rz / np.dot(p, q)
            x += alpha * p
            r -= alpha * q
            z = self.dinv * r
            rz_new = np.dot(r, z)
            p = z + (rz_new / rz) * p
            rz = rz_new
            it += 1
        self.iters += it
        self.u = x
        return x
