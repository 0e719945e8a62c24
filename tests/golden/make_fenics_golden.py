#!/usr/bin/env python
"""Pin the oracle (and through it the CUDA path) against the REAL reference: run the unmodified functions of
ziyu0425/PDE-Solver's fenics_mcp_server.py under FEniCS/DOLFIN 2019.1.0 and record what they produce.

    # inside the reference's own environment (Dockerfile: conda-forge fenics-dolfin=2019.1.0)
    python tests/golden/make_fenics_golden.py /path/to/PDE-Solver  [--out tests/golden/fenics_vectors.npz]

FEniCS is not installable in this repository's build container (no conda, no wheels, no network), so the output file is
NOT committed yet: `parity unpinned` in DESIGN.md stays until somebody runs this script once and commits the .npz.
tests/test_fenics_golden.py activates by itself when the file exists and compares (a) the oracle and (b), on a GPU, the
CUDA path against it: meshes / cells / dof maps / Dirichlet sets bit-exact, nodal solutions <= 1e-8 relative L2.

What is recorded, for reorder_dofs_serial False and True (DOLFIN's default):
  per mesh      : mesh.coordinates(), mesh.cells(), dof_to_vertex_map(V), V.tabulate_dof_coordinates(), the cell-dof table,
                  the DirichletBC dof sets of the predicates the reference uses (:233-241, 373-376, 606-628, 1531-1534)
  per case      : the TimeSeriesField the reference's _solve_* function returns (coords, values, times) - these functions
                  pair tabulate_dof_coordinates() with get_local() (:535, 649, 663, 1865-1866), so they are numbering-proof
Cases: BASELINE config 1 as is; configs 2-5 at reduced sizes that the reference's sparse LU finishes in seconds.
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np

CASES = {
    # name: (function, kwargs)  - the same kwargs tests/test_fenics_golden.py replays
    "cfg1_heat1d": ("_solve_heat_1d_raw", dict(length=2.0, nx=100, diffusivity=1.0, T_left=20.0, T_right=0.0, T_initial=0.0,
                                               dt=0.01, num_steps=200)),
    "cfg2_heat2d_r": ("_solve_heat_2d_raw", dict(Lx=1.0, Ly=1.0, nx=48, ny=40, diffusivity=1.0, T_boundary=0.0,
                                                 T_initial=20.0, dt=0.01, num_steps=10)),
    "cfg3_elast3d_r": ("_solve_elasticity_3d_static", dict(Lx=1.0, Ly=0.2, Lz=0.2, nx=40, ny=8, nz=8, E=210e9, nu=0.3,
                                                           body_fx=0.0, body_fy=0.0, body_fz=-76518.0, quantity="stress")),
    "cfg4_heat3d_r": ("_solve_heat_3d_raw", dict(Lx=1.0, Ly=1.0, Lz=1.0, nx=16, ny=16, nz=16, diffusivity=1.0, T_boundary=0.0,
                                                 T_initial=20.0, dt=0.01, num_steps=5)),
    "cfg5_elast3d_r": ("_solve_elasticity_3d_static", dict(Lx=1.0, Ly=0.2, Lz=0.2, nx=80, ny=16, nz=16, E=210e9, nu=0.3,
                                                           body_fx=0.0, body_fy=0.0, body_fz=-76518.0, quantity="stress")),
    "heat3d_directional": ("_solve_heat_3d_raw", dict(Lx=1.0, Ly=0.5, Lz=0.5, nx=12, ny=6, nz=6, diffusivity=0.7,
                                                      T_boundary=0.0, T_initial=5.0, dt=0.02, num_steps=4, T_left=100.0,
                                                      T_right=20.0, T_side=10.0)),
    "heat2d_cosine": ("_solve_heat_2d_raw", dict(Lx=1.0, Ly=1.0, nx=16, ny=16, diffusivity=1.0, T_boundary=0.0, T_initial=0.0,
                                                 dt=0.01, num_steps=3, initial_type="cosine", initial_amplitude=2.0,
                                                 initial_wavenumber=3.0)),
    "elast2d_strain": ("_solve_elasticity_2d_static", dict(Lx=1.0, Ly=0.5, nx=24, ny=12, E=210e9, nu=0.3, body_fx=0.0,
                                                           body_fy=-76518.0, quantity="strain", plane_stress=True)),
}
MESHES = {"interval_100": (1, [100], [2.0]), "rect_5x3": (2, [5, 3], [1.0, 0.6]), "box_4x3x2": (3, [4, 3, 2], [1.0, 0.6, 0.35])}


def load_reference(root):
    """Import the reference's tool file without starting its MCP server (FastMCP may be absent: stub it)."""
    if "mcp" not in sys.modules:
        try:
            import mcp.server.fastmcp  # noqa: F401
        except Exception:
            m = types.ModuleType("mcp")
            ms = types.ModuleType("mcp.server")
            mf = types.ModuleType("mcp.server.fastmcp")

            class FastMCP:                      # decorators only
                def __init__(self, *a, **k):
                    pass

                def tool(self, *a, **k):
                    return lambda f: f

                def run(self, *a, **k):
                    pass
            mf.FastMCP = FastMCP
            sys.modules.update({"mcp": m, "mcp.server": ms, "mcp.server.fastmcp": mf})
    spec = importlib.util.spec_from_file_location("ref_fenics_mcp_server", os.path.join(root, "fenics_mcp_server.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def mesh_records(out, tag):
    import dolfin as df
    for name, (dim, n, L) in MESHES.items():
        if dim == 1:
            mesh = df.IntervalMesh(n[0], 0.0, L[0])
        elif dim == 2:
            mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(L[0], L[1]), n[0], n[1])
        else:
            mesh = df.BoxMesh(df.Point(0.0, 0.0, 0.0), df.Point(*L), *n)
        V = df.FunctionSpace(mesh, "P", 1)
        k = f"{tag}/{name}"
        out[f"{k}/coordinates"] = mesh.coordinates().copy()
        out[f"{k}/cells"] = mesh.cells().astype(np.int32)
        out[f"{k}/dof_to_vertex_map"] = np.asarray(df.dof_to_vertex_map(V), dtype=np.int64)
        out[f"{k}/dof_coordinates"] = V.tabulate_dof_coordinates().copy()
        out[f"{k}/cell_dofs"] = np.array([V.dofmap().cell_dofs(c) for c in range(mesh.num_cells())], dtype=np.int32)
        tol = 1e-14
        preds = {"left": lambda x, on: on and df.near(x[0], 0.0, tol), "all": lambda x, on: on}
        if dim == 3:
            preds["other_faces"] = lambda x, on: on and not df.near(x[0], 0.0, tol) and not df.near(x[0], L[0], tol)
        for pn, pred in preds.items():
            bc = df.DirichletBC(V, df.Constant(1.0), pred)
            out[f"{k}/bc_{pn}"] = np.array(sorted(bc.get_boundary_values().keys()), dtype=np.int64)
        if dim > 1:
            W = df.VectorFunctionSpace(mesh, "P", 1)
            out[f"{k}/vector_cell_dofs"] = np.array([W.dofmap().cell_dofs(c) for c in range(mesh.num_cells())], dtype=np.int32)
            out[f"{k}/vector_dof_coordinates"] = W.tabulate_dof_coordinates().copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference_root")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "fenics_vectors.npz"))
    args = ap.parse_args()
    import dolfin as df
    ref = load_reference(args.reference_root)
    out = {"dolfin_version": np.array(df.__version__)}
    for reorder in (False, True):
        df.parameters["reorder_dofs_serial"] = reorder
        tag = f"reorder_{int(reorder)}"
        mesh_records(out, tag)
        for name, (fn, kw) in CASES.items():
            f = getattr(ref, fn)(**kw)
            out[f"{tag}/{name}/coords"] = np.asarray(f.coords, dtype=np.float64)
            out[f"{tag}/{name}/values"] = np.asarray(f.values, dtype=np.float64)
            out[f"{tag}/{name}/times"] = np.asarray(f.times, dtype=np.float64)
            print(f"{tag} {name}: {out[f'{tag}/{name}/values'].shape}", flush=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
