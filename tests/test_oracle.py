"""Known-answer tests that pin the CPU oracle; no GPU, no reference."""
import numpy as np
import pytest

from oracle import fem_oracle as fo


def test_mesh_counts_and_order():
    m = fo.box_mesh((0, 0, 0), (1, 0.2, 0.2), 4, 3, 2)
    assert m.coords.shape == (5 * 4 * 3, 3) and m.cells.shape == (6 * 24, 4)
    # vertex id = iz*(nx+1)*(ny+1) + iy*(nx+1) + ix, x fastest
    assert np.array_equal(m.coords[7], [0.5, 0.2 * 1 / 3 if False else (1 * 0.2) / 3, 0.0])
    # first grid cell, generation order, DOLFIN BoxMesh tets
    v = [0, 1, 5, 6, 20, 21, 25, 26]
    expect = [[v[a] for a in t] for t in fo.BOX_TETS]
    assert m.cells_raw[:6].tolist() == expect
    assert np.all(np.diff(m.cells, axis=1) > 0)
    vol, _ = fo._gradients(m)
    assert np.allclose(vol.sum(), 1 * 0.2 * 0.2) and np.allclose(vol, vol[0])
    r = fo.rectangle_mesh(0, 0, 1, 1, 3, 2)
    assert r.cells_raw[:2].tolist() == [[0, 1, 5], [0, 4, 5]]
    i = fo.interval_mesh(100, 0.0, 2.0)
    assert i.coords[-1, 0] == 2.0 and i.coords[37, 0] == 0.0 + (2.0 / 100) * 37


def test_coordinate_expressions_differ_2d_vs_3d():
    # A.1: a + ((b-a)/n)*i (2D) vs a + (i*(b-a))/n (3D) differ in the last bit for non-dyadic n
    r = fo.rectangle_mesh(0, 0, 0.2, 0.2, 10, 10).coords[:11, 0]
    b = fo.box_mesh((0, 0, 0), (0.2, 0.2, 0.2), 10, 1, 1).coords[:11, 0]
    assert np.max(np.abs(r - b)) < 1e-16 and np.any(r != b)
    assert fo.box_mesh((0, 0, 0), (1, 0.2, 0.2), 320, 1, 1).coords[320, 0] == 1.0


def test_stencil_facts_3d():
    # (vi): interior rows of K are the 7-point Laplacian, M has the 15-point Kuhn weights
    n, h = 4, 0.25
    m = fo.box_mesh((0, 0, 0), (1, 1, 1), n, n, n)
    K, M = fo.assemble_stiffness_mass(m)
    c = 2 + 5 * 2 + 25 * 2
    rowK = K[c].toarray().ravel()
    rowM = M[c].toarray().ravel() / h ** 3
    off = lambda dx, dy, dz: c + dx + 5 * dy + 25 * dz
    assert np.isclose(rowK[c], 6 * h) and np.isclose(rowK[off(1, 0, 0)], -h)
    assert np.isclose(rowK[off(1, 1, 0)], 0) and np.isclose(rowK[off(1, 1, 1)], 0)
    assert np.isclose(rowM[c], 0.4) and np.isclose(rowM[off(1, 0, 0)], 0.05)
    assert np.isclose(rowM[off(1, 1, 0)], 1 / 30) and np.isclose(rowM[off(-1, -1, -1)], 0.05)
    assert np.isclose(rowM[off(1, -1, 0)], 0) and np.isclose(rowM.sum(), 1.0)
    assert np.count_nonzero(np.abs(rowM) > 1e-14) == 15


def test_stencil_facts_2d():
    m = fo.rectangle_mesh(0, 0, 1, 1, 4, 4)
    K, M = fo.assemble_stiffness_mass(m)
    c = 12
    rk, rm = K[c].toarray().ravel(), M[c].toarray().ravel() * 16
    assert np.isclose(rk[c], 4) and np.isclose(rk[c + 1], -1) and np.isclose(rk[c + 6], 0)
    assert np.isclose(rm[c], 0.5) and np.isclose(rm[c + 6], 1 / 12) and np.isclose(rm[c + 4], 0)


def test_heat_1d_steady_linear():
    f = fo.solve_heat(1, [2.0], [100], 1.0, T_left=20.0, T_right=0.0, steady=True)
    x = f.coords[:, 0]
    assert np.allclose(f.values[0], 20.0 * (1 - x / 2.0), atol=1e-11)


def test_heat_1d_sine_decay_closed_form():
    # (iv-a): sin(jπx/L) is an eigenvector of M and K in 1D => exact per-step decay factor
    L, nx, dt, kappa, j = 2.0, 64, 0.01, 1.0, 3
    h = L / nx
    mesh = fo.interval_mesh(nx, 0, L)
    K, M = fo.assemble_stiffness_mass(mesh)
    x = mesh.coords[:, 0]
    u0 = np.sin(j * np.pi * x / L)
    th = j * np.pi * h / L
    lamM, lamK = h / 6 * (4 + 2 * np.cos(th)), (2 - 2 * np.cos(th)) / h
    A, _ = fo.apply_bc_rowwise((M + dt * kappa * K).tocsr(), np.zeros(nx + 1), np.array([0, nx]), np.zeros(2))
    b = M @ u0
    b[[0, nx]] = 0
    u1 = fo.lu_solve(A, b)
    assert np.allclose(u1[1:-1], lamM / (lamM + dt * kappa * lamK) * u0[1:-1], atol=1e-13)


def test_heat_rowwise_equals_symmetric():
    a = fo.solve_heat(3, [1, 1, 1], [6, 5, 4], 1.0, T_initial=20.0, num_steps=3, T_boundary=2.0)
    b = fo.solve_heat(3, [1, 1, 1], [6, 5, 4], 1.0, T_initial=20.0, num_steps=3, T_boundary=2.0, symmetric=True)
    assert fo.rel_l2(a.values, b.values) < 1e-13
    assert a.values.shape == (4, 7 * 6 * 5) and a.times[-1] == 3 * 0.01


def test_heat_boundary_sets():
    f = fo.solve_heat(3, [1, 1, 1], [4, 3, 2], 1.0, num_steps=0)
    assert f.aux["bc_dofs"].size == 5 * 4 * 3 - 3 * 2 * 1
    # directional: other_faces excludes vertices on x=0 / x=Lx (topological facet rule)
    g = fo.solve_heat(3, [1, 1, 1], [4, 3, 2], 1.0, num_steps=0, T_side=5.0)
    ix = g.aux["bc_dofs"] % 5
    assert ix.min() == 1 and ix.max() == 3
    h = fo.solve_heat(3, [1, 1, 1], [2, 3, 2], 1.0, num_steps=0, T_side=5.0)
    assert h.aux["bc_dofs"].size == 0      # nx=2: every side facet touches an x-end


def test_bar_1d_nodal_exact():
    # (ii): u = f/(EA) (L x - x²/2) exactly at the nodes
    L, nx, E, A, f = 1.5, 40, 210e9, 2.0, 1e6
    r = fo.solve_elasticity(1, [L], [nx], E, body=[f], quantity="strain", area=A)
    x = r.coords[:, 0]
    assert np.allclose(r.aux["u"][:, 0], f / (E * A) * (L * x - x * x / 2), rtol=1e-10, atol=1e-17)
    # cell strain is exact at cell midpoints; projection reproduces it in the interior to O(h²)
    assert np.allclose(r.values[0][5:-5], f / (E * A) * (L - x[5:-5]), rtol=1e-3)


@pytest.mark.parametrize("dim", [2, 3])
def test_zero_body_force_zero_von_mises(dim):
    # (iii) docstring fenics_mcp_server.py:2639-2642
    r = fo.solve_elasticity(dim, [1, 0.5, 0.5][:dim], [6, 4, 3][:dim], 210e9, 0.3)
    assert np.all(r.values == 0.0)


def test_patch_test_linear_displacement():
    # (v): a linear displacement field has zero residual at nodes whose patch is complete
    m = fo.box_mesh((0, 0, 0), (1, 0.5, 0.25), 4, 4, 4)
    lam, mu = fo.lame(210e9, 0.3, 3)
    A = fo.assemble_elasticity(m, lam, mu)
    G = np.array([[1e-3, 2e-4, 0], [-1e-4, 5e-4, 3e-4], [2e-4, 0, -4e-4]])
    u = (m.coords @ G.T).ravel()
    r = (A @ u).reshape(-1, 3)
    ijk = np.stack(np.unravel_index(np.arange(m.nv), (5, 5, 5)), axis=1)
    interior = np.all((ijk > 0) & (ijk < 4), axis=1)
    assert np.abs(r[interior]).max() < 1e-12 * A.diagonal().max() * 1e-3
    vm = fo.von_mises_cells(m, u.reshape(-1, 3), lam, mu, "stress")
    assert np.allclose(vm, vm[0], rtol=1e-10)


def test_cantilever_sane():
    # A.7: 40x8x8 beam tip deflection ≈ -1.307e-5 (Euler-Bernoulli -1.366e-5)
    r = fo.solve_elasticity(3, [1, 0.2, 0.2], [40, 8, 8], 210e9, 0.3, body=[0, 0, -76518.0])
    uz = r.aux["u"][:, 2]
    assert abs(uz.min() / -1.307239e-5 - 1) < 1e-5
    assert abs(r.values.max() / 1.12021e6 - 1) < 1e-4    # L2 projection may undershoot below 0


def test_project_p2_reproduces_quadratic_moments():
    # P2 interpolation of a quadratic is exact => projection == L2 projection of the function
    m = fo.rectangle_mesh(0, 0, 1, 1, 6, 5)
    fn = lambda x: 1 + x[:, 0] ** 2 - x[:, 0] * x[:, 1]
    lin = lambda x: 2 - x[:, 0] + 3 * x[:, 1]
    assert np.allclose(fo.project_p2_expression(m, lin), lin(m.coords), atol=1e-12)
    u = fo.project_p2_expression(m, fn)
    _, M = fo.assemble_stiffness_mass(m)
    assert np.isclose((M @ u).sum(), 1 + 1 / 3 - 1 / 4)


def test_weighted_forms_known_answers():
    # curvilinear tools: weighted mass sums to the integral of the weight (exact for polynomial weights) and the
    # steady radial solutions converge to the analytic log / 1/r profiles
    m = fo.interval_mesh(10, 0.5, 2.0)
    _, M, l = fo.assemble_weighted(m, lambda x: x[:, 0], 1)
    assert np.isclose(M.sum(), (2.0 ** 2 - 0.5 ** 2) / 2) and np.isclose(l.sum(), M.sum())
    _, M, l = fo.assemble_weighted(m, lambda x: x[:, 0] ** 2, 2)
    assert np.isclose(M.sum(), (2.0 ** 3 - 0.5 ** 3) / 3) and np.isclose(l.sum(), M.sum())
    b = fo.box_mesh((0.5, 0, 0), (2.0, 1.0, 1.0), 3, 2, 2)
    K, M, l = fo.assemble_weighted(b, lambda x: x[:, 0] ** 2, 2)
    assert np.isclose(M.sum(), (8 - 0.125) / 3) and np.allclose(K @ np.ones(b.nv), 0, atol=1e-12)
    K1, M1 = fo.assemble_stiffness_mass(b)
    Kc, Mc, _ = fo.assemble_weighted(b, lambda x: np.full(x.shape[0], 3.0), 1)      # constant weight = plain forms
    assert abs(Kc - 3 * K1).max() < 1e-12 and abs(Mc - 3 * M1).max() < 1e-14
    errs = []
    for n in (100, 200):
        f = fo.solve_heat_curvilinear("1d_cylindrical", 0.5, 2.0, [n], 1.0, steady=True, T_inner=100.0, T_outer=20.0)
        r = f.coords[:, 0]
        errs.append(np.abs(f.values[0] - (100 - 80 * np.log(r / 0.5) / np.log(4.0))).max())
    assert errs[1] < errs[0] / 3.5 and errs[1] < 1e-3                                   # O(h^2)
    f = fo.solve_heat_curvilinear("1d_spherical", 0.5, 2.0, [400], 1.0, steady=True, T_inner=100.0, T_outer=20.0)
    r = f.coords[:, 0]
    assert np.abs(f.values[0] - (20 + 80 * (1 / r - 0.5) / 1.5)).max() < 1e-3


# ---------------------------------------------------------------- cylinder / composite-core branches of solve_heat_3D
def test_special_3d_reduces_to_box_solver():
    """solve_heat_3d_special with unit weight and a core of the same diffusivity is the plain box solve."""
    kw = dict(T_boundary=1.0, T_initial=4.0, dt=0.02, num_steps=3, source_type="constant", source_value=2.0)
    ref = fo.solve_heat(3, [1.0, 0.5, 0.4], [5, 4, 3], 0.7, **kw)
    a = fo.solve_heat_3d_special(1.0, 0.5, 0.4, [5, 4, 3], 0.7, core_radius=0.3, core_diffusivity=0.7, **kw)
    assert np.allclose(a.values, ref.values, rtol=0, atol=1e-12 * np.abs(ref.values).max())
    assert np.array_equal(a.coords, ref.coords)


def test_special_3d_cylinder_mesh_and_side_set():
    """The BoxMesh 'cylinder' (:527-529): box [0,Lx]x[-R,R]^2 with int(ny*2R) cells; side_boundary_cylinder finds no
    facet (no boundary facet of the box has all its vertices at r == R), so T_side alone constrains nothing and the
    first step from a constant state with no source leaves it constant."""
    R = 0.3
    f = fo.solve_heat_3d_special(1.0, 9.0, 9.0, [4, 10, 10], 1.0, T_initial=5.0, dt=0.01, num_steps=2,
                                 geometry_type="cylinder", cylinder_radius=R, T_side=100.0)
    assert f.coords.shape[0] == 5 * 7 * 7
    assert f.coords[:, 1].min() == -R and f.coords[:, 2].max() == R
    assert np.allclose(f.values, 5.0, rtol=0, atol=1e-10)


def test_special_3d_core_marks_cells_by_vertices_and_midpoint():
    m = fo.box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 2, 4, 4)
    X = m.coords[m.cells]
    rv = np.sqrt(X[:, :, 1] ** 2 + X[:, :, 2] ** 2)
    inside = (rv < 0.6).all(axis=1)
    # quarter disc of radius 0.6 on the 0.25 lattice: (y,z) lattice points (j,k) with j^2+k^2 < 5.76 are j,k in {0,1,2}
    # except (2,2).  Footprint cells (0,0),(0,1),(1,0) lie inside with all 6 tets; every Kuhn tet of cell (1,1) contains
    # the far corner (2,2), so none of its tets is marked: 3 cells x 6 tets x 2 cells along x
    assert inside.sum() == 36
    f0 = fo.solve_heat_3d_special(1.0, 1.0, 1.0, [2, 4, 4], 1.0, T_initial=3.0, dt=0.05, num_steps=2,
                                  core_radius=0.6, core_diffusivity=50.0)
    f1 = fo.solve_heat_3d_special(1.0, 1.0, 1.0, [2, 4, 4], 1.0, T_initial=3.0, dt=0.05, num_steps=2)
    assert np.abs(f0.values[-1] - f1.values[-1]).max() > 1e-3      # the core changes the solution


# ---------------------------------------------------------------- (iv-b) Fourier symbol of the interior stencil
@pytest.mark.parametrize("dim,n,L", [(3, [6, 5, 7], [1.0, 0.6, 0.35]), (2, [7, 6], [1.0, 0.4]), (3, [4, 4, 4], [1, 1, 1])])
def test_assembled_interior_rows_have_the_closed_form_symbol(dim, n, L):
    """cos(theta.n + phi) is an eigenvector of the interior rows of alpha*M + beta*K with eigenvalue
    sum_d w_d cos(theta.d), w_d the closed-form Kuhn / right-diagonal P1 weights (mass 2/5, 1/20, 1/30,
    1/20 and the 7-point Laplacian in 3-D; 1/2, 1/12 and the 5-point Laplacian in 2-D)."""
    alpha, beta = 1.0, 0.01
    theta = np.array([0.31, 0.73, 1.17][:dim])
    h = [a / b for a, b in zip(L, n)]
    if dim == 3:
        hx, hy, hz = h
        vol, kx, ky, kz = hx * hy * hz, hy * hz / hx, hx * hz / hy, hx * hy / hz
        w = {(0, 0, 0): alpha * vol * 2 / 5 + beta * 2 * (kx + ky + kz), (1, 0, 0): alpha * vol / 20 - beta * kx,
             (0, 1, 0): alpha * vol / 20 - beta * ky, (0, 0, 1): alpha * vol / 20 - beta * kz,
             (1, 1, 0): alpha * vol / 30, (1, 0, 1): alpha * vol / 30, (0, 1, 1): alpha * vol / 30,
             (1, 1, 1): alpha * vol / 20}
    else:
        hx, hy = h
        vol = hx * hy
        w = {(0, 0): alpha * vol / 2 + beta * 2 * (hy / hx + hx / hy), (1, 0): alpha * vol / 12 - beta * hy / hx,
             (0, 1): alpha * vol / 12 - beta * hx / hy, (1, 1): alpha * vol / 12}
    zero = tuple([0] * dim)
    C = w[zero] + sum(2 * v * np.cos(np.dot(theta, d)) for d, v in w.items() if d != zero)
    m = fo.make_mesh(dim, L, n)
    K, M = fo.assemble_stiffness_mass(m)
    A = (alpha * M + beta * K).tocsr()
    nn = [k + 1 for k in n]
    idx = np.indices(nn[::-1]).reshape(dim, -1)[::-1]           # natural order, x fastest
    x = np.cos(0.4 + sum(theta[d] * idx[d] for d in range(dim)))
    interior = np.all([(idx[d] > 0) & (idx[d] < nn[d] - 1) for d in range(dim)], axis=0)
    assert np.abs((A @ x)[interior] - C * x[interior]).max() <= 1e-14 * max(abs(v) for v in w.values())
