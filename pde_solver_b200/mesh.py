"""Structured simplicial meshes, P1 dof maps and Dirichlet sets, generated on the GPU.

Mirrors what the reference obtains from DOLFIN (fenics_mcp_server.py:229-230, 369-370, 533-535,
1649-1650, 1804-1805, 233-241, 373-376, 606-628): IntervalMesh / RectangleMesh("right") / BoxMesh
vertex coordinates and connectivity, FunctionSpace / VectorFunctionSpace cell-dof tables, and the
topological DirichletBC vertex sets.  Numbering is DOLFIN's with reorder_dofs_serial=False; DOLFIN's default
(reorder_dofs_serial=True: a graph reordering of the dofs that cannot be derived from anything in the reference) can be
plugged in as a permutation, see set_dof_permutation."""
import ctypes as C

import numpy as np

from . import _lib


def _n3(n):
    return list(n) + [0] * (3 - len(n))


# ---- pluggable dof numbering ---------------------------------------------------------------------
# vertex_of_dof[d] = natural (x fastest) vertex index that scalar dof d sits on; for DOLFIN this is
# dof_to_vertex_map(FunctionSpace(mesh, "P", 1)) - tests/golden/make_fenics_golden.py records it for both settings of
# reorder_dofs_serial.  Registered per mesh; every exported array (coords, values, cell-dof tables, Dirichlet masks) of
# that mesh then comes out in the plugged numbering.  Nothing on the device changes: the solver keeps natural order.
_dof_perm = {}


def _key(dim, n):
    return (int(dim),) + tuple(int(v) for v in list(n)[:dim])


def set_dof_permutation(dim, n, vertex_of_dof):
    """Register (or, with None, remove) the dof numbering of the P1 space on the mesh (dim, n)."""
    if vertex_of_dof is None:
        _dof_perm.pop(_key(dim, n), None)
        return
    v = np.asarray(vertex_of_dof, dtype=np.int64)
    nv, _ = _lib.mesh_counts(dim, n)
    if v.shape != (nv,) or not np.array_equal(np.sort(v), np.arange(nv)):
        raise ValueError("vertex_of_dof must be a permutation of the mesh vertices")
    _dof_perm[_key(dim, n)] = v


def dof_permutation(dim, n):
    """vertex_of_dof of the mesh, or None for the natural numbering."""
    return _dof_perm.get(_key(dim, n))


def to_dof_order(dim, n, a, axis=-1):
    """Reorder the vertex axis of `a` (natural order) into the plugged dof numbering."""
    v = dof_permutation(dim, n)
    return a if v is None else np.take(a, v, axis=axis)


def coordinates(dim, n, L, ctx=None):
    """(nverts, dim) float64, bit-exact DOLFIN expressions."""
    ctx = ctx or _lib.default_context()
    nv, _ = _lib.mesh_counts(dim, n)
    out = np.empty((nv, dim), dtype=np.float64)
    _lib.check(_lib.lib().pde_mesh_coords(ctx.handle, int(dim), _lib.i3(n), _lib.d3(L), _lib.ptr(out)))
    return out


def coordinates_box(dim, n, lo, hi, ctx=None):
    """(nverts, dim) float64 on [lo, hi]: IntervalMesh(n, a, b) / RectangleMesh(Point(a..), Point(b..)) / BoxMesh."""
    ctx = ctx or _lib.default_context()
    nv, _ = _lib.mesh_counts(dim, n)
    out = np.empty((nv, dim), dtype=np.float64)
    _lib.check(_lib.lib().pde_mesh_coords_box(ctx.handle, int(dim), _lib.i3(n), _lib.d3(lo, 0.0), _lib.d3(hi, 0.0),
                                              _lib.ptr(out)))
    return out


def cells(dim, n, ordered=True, ctx=None):
    """(ncells, dim+1) int32 connectivity; ordered=True is the state after mesh.order()."""
    ctx = ctx or _lib.default_context()
    _, nc = _lib.mesh_counts(dim, n)
    out = np.empty((nc, dim + 1), dtype=np.int32)
    _lib.check(_lib.lib().pde_mesh_cells(ctx.handle, int(dim), _lib.i3(n), 1 if ordered else 0, _lib.ptr(out)))
    return out


def cell_dofs(dim, n, ncomp=1, layout="blocked", ctx=None):
    """(ncells, ncomp*(dim+1)) int32 P1 cell-dof table; layout 'blocked' (UFC) or 'interleaved'."""
    ctx = ctx or _lib.default_context()
    _, nc = _lib.mesh_counts(dim, n)
    out = np.empty((nc, ncomp * (dim + 1)), dtype=np.int32)
    _lib.check(_lib.lib().pde_dofmap_cells(ctx.handle, int(dim), _lib.i3(n), int(ncomp),
                                            0 if layout == "blocked" else 1, _lib.ptr(out)))
    v = dof_permutation(dim, n)
    if v is not None:       # natural dof (c, vertex) -> plugged scalar numbering of its vertex, same layout
        nv = v.size
        dof_of_vertex = np.empty(nv, dtype=np.int64)
        dof_of_vertex[v] = np.arange(nv)
        if layout == "blocked":
            comp, vert = out // nv, out % nv
            out = (comp * nv + dof_of_vertex[vert]).astype(np.int32)
        else:
            vert, comp = out // ncomp, out % ncomp
            out = (dof_of_vertex[vert] * ncomp + comp).astype(np.int32)
    return out


def dirichlet(dim, n, bc, ctx=None):
    """(mask uint8 [nverts], values float64 [nverts]) of a pde_bc specification."""
    ctx = ctx or _lib.default_context()
    nv, _ = _lib.mesh_counts(dim, n)
    mask = np.empty(nv, dtype=np.uint8)
    vals = np.empty(nv, dtype=np.float64)
    _lib.check(_lib.lib().pde_boundary_mask(ctx.handle, int(dim), _lib.i3(n), C.byref(bc), _lib.ptr(mask),
                                             _lib.ptr(vals)))
    return to_dof_order(dim, n, mask), to_dof_order(dim, n, vals)


def heat_bc(dim, T_boundary=0.0, T_left=None, T_right=None, T_side=None):
    """The reference's heat DirichletBC lists as a pde_bc (1D: left/right :233-241; 2D: all :373-376;
    3D: all, or directional left/right/other_faces :576-628)."""
    if dim == 1:
        return _lib.make_bc({0: T_left, 1: T_right})
    if dim == 3 and (T_left is not None or T_right is not None or T_side is not None):
        faces = {}
        if T_left is not None:
            faces[0] = T_left
        if T_right is not None:
            faces[1] = T_right
        if T_side is not None:
            for f in (2, 3, 4, 5):
                faces[f] = T_side
        return _lib.make_bc(faces, side_excludes_xends=True)
    return _lib.make_bc({f: T_boundary for f in range(2 * dim)})
