"""Committed golden vectors (tests/golden/oracle_vectors.npz, made by tests/golden/make_golden_vectors.py).

CPU: the oracle keeps reproducing them (integer arrays and coordinates bit-exact, solves to 1e-12).
GPU: the CUDA path matches them (bit-exact for meshes / dof maps / boundary sets, <= 1e-8 rel-L2 for solves)."""
import os

import numpy as np
import pytest

from oracle import fem_oracle as fo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
MESHES = [("m1", 1, [7], [0.3]), ("m2", 2, [5, 3], [1.0, 0.6]), ("m3", 3, [4, 3, 2], [1.0, 0.2, 0.2])]
TOL = 1e-8


@pytest.mark.parametrize("tag,dim,n,L", MESHES)
def test_oracle_meshes_reproduce_golden(tag, dim, n, L):
    m = fo.make_mesh(dim, L, n)
    assert np.array_equal(m.coords, G[f"{tag}_coords"])
    assert np.array_equal(m.cells, G[f"{tag}_cells"]) and np.array_equal(m.cells_raw, G[f"{tag}_cells_raw"])
    assert np.array_equal(fo.dirichlet_dofs(m, lambda x, ob: np.ones(x.shape[0], bool)), G[f"{tag}_boundary"])


def test_oracle_solves_reproduce_golden():
    f = fo.solve_heat(1, [2.0], [100], 1.0, T_initial=0.0, dt=0.01, num_steps=200, T_left=20.0, T_right=0.0)
    assert fo.rel_l2(f.values[[0, 1, 10, 100, 200]], G["cfg1_values"]) <= 1e-12
    f = fo.solve_heat(3, [1, 1, 1], [16, 16, 16], 1.0, T_initial=20.0, dt=0.01, num_steps=5, T_boundary=0.0)
    assert fo.rel_l2(f.values[[1, 5]], G["heat3d_values"]) <= 1e-12
    g = fo.solve_elasticity(3, [1, 0.2, 0.2], [20, 4, 4], 210e9, 0.3, body=[0, 0, -76518.0], quantity="stress")
    assert fo.rel_l2(g.values[0], G["cantilever_stress"]) <= 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dim,n,L", MESHES)
def test_gpu_meshes_match_golden(tag, dim, n, L):
    import pde_solver_b200 as P
    assert np.array_equal(P.mesh.coordinates(dim, n, L), G[f"{tag}_coords"])
    assert np.array_equal(P.mesh.cells(dim, n, ordered=True), G[f"{tag}_cells"])
    assert np.array_equal(P.mesh.cells(dim, n, ordered=False), G[f"{tag}_cells_raw"])
    if dim > 1:
        assert np.array_equal(P.mesh.cell_dofs(dim, n, dim, "interleaved"), G[f"{tag}_vdofs_interleaved"])
    mask, _ = P.mesh.dirichlet(dim, n, P._lib.make_bc({f: 0.0 for f in range(2 * dim)}))
    assert np.array_equal(np.nonzero(mask)[0], G[f"{tag}_boundary"])


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["jacobi", "gmg"])
def test_gpu_solves_match_golden(precond):
    import pde_solver_b200 as P
    f = P._solve_heat_1d_raw(2.0, 100, 1.0, 20.0, 0.0, 0.0, 0.01, 200, precond=precond, as_arrays=True)
    assert fo.rel_l2(f.values[[0, 1, 10, 100, 200]], G["cfg1_values"]) <= TOL
    f = P._solve_heat_2d_raw(1.0, 1.0, 32, 32, 1.0, 0.0, 20.0, 0.01, 10, precond=precond, as_arrays=True)
    assert fo.rel_l2(f.values[[1, 10]], G["heat2d_values"]) <= TOL
    f = P._solve_heat_3d_raw(1, 1, 1, 16, 16, 16, 1.0, 0.0, 20.0, 0.01, 5, precond=precond, as_arrays=True)
    assert fo.rel_l2(f.values[[1, 5]], G["heat3d_values"]) <= TOL
    for q in ("stress", "strain"):
        g = P._solve_elasticity_3d_static(1, 0.2, 0.2, 20, 4, 4, 210e9, 0.3, 0.0, 0.0, -76518.0, q, precond=precond,
                                          as_arrays=True)
        assert fo.rel_l2(g.values[0], G[f"cantilever_{q}"]) <= TOL
