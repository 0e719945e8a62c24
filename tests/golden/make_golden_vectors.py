"""Generate tests/golden/oracle_vectors.npz: small input/output vectors of the CPU oracle.

The reference itself (FEniCS/DOLFIN 2019.1.0) cannot run in this container, so these vectors pin the
ORACLE (a regression fixture: oracle/fem_oracle.py must keep reproducing them bit-for-bit in the integer
arrays and to 1e-13 in the solves) and give the GPU tests a committed target that does not depend on
SciPy's LU at run time.  They are not reference outputs: parity stays "unpinned" (oracle header)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fem_oracle as fo  # noqa: E402

out = {}
# meshes / dof maps / boundary sets (bit-exact targets)
for tag, dim, n, L in (("m1", 1, [7], [0.3]), ("m2", 2, [5, 3], [1.0, 0.6]), ("m3", 3, [4, 3, 2], [1.0, 0.2, 0.2])):
    m = fo.make_mesh(dim, L, n)
    out[f"{tag}_coords"] = m.coords
    out[f"{tag}_cells"] = m.cells.astype(np.int32)
    out[f"{tag}_cells_raw"] = m.cells_raw.astype(np.int32)
    if dim > 1:
        out[f"{tag}_vdofs_interleaved"] = fo.cell_dofs_vector(m, dim, "interleaved").astype(np.int32)
    bnd = fo.dirichlet_dofs(m, lambda x, ob: np.ones(x.shape[0], bool))
    out[f"{tag}_boundary"] = bnd.astype(np.int64)
# BASELINE config 1: 1D rod, 100 cells, 20/0 Dirichlet, 200 backward-Euler steps
f = fo.solve_heat(1, [2.0], [100], 1.0, T_initial=0.0, dt=0.01, num_steps=200, T_left=20.0, T_right=0.0)
out["cfg1_values"] = f.values[[0, 1, 10, 100, 200]]
# 2D / 3D heat, reduced sizes of configs 2 and 4
f = fo.solve_heat(2, [1.0, 1.0], [32, 32], 1.0, T_initial=20.0, dt=0.01, num_steps=10, T_boundary=0.0)
out["heat2d_values"] = f.values[[1, 10]]
f = fo.solve_heat(3, [1, 1, 1], [16, 16, 16], 1.0, T_initial=20.0, dt=0.01, num_steps=5, T_boundary=0.0)
out["heat3d_values"] = f.values[[1, 5]]
# cantilever (configs 3/5 at reduced size): projected von Mises stress and strain, displacement
for q in ("stress", "strain"):
    g = fo.solve_elasticity(3, [1, 0.2, 0.2], [20, 4, 4], 210e9, 0.3, body=[0, 0, -76518.0], quantity=q)
    out[f"cantilever_{q}"] = g.values[0]
out["cantilever_u"] = g.aux["u"]
path = os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes", len(out), "arrays")
