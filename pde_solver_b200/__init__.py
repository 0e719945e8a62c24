"""B200-native structured-mesh P1 heat / linear-elasticity solver.

Drop-in for the FEniCS/PETSc path of ziyu0425/PDE-Solver (`fenics_mcp_server.py`): the six
`_solve_*` functions of `solvers` mirror the reference's raw solvers (same names, argument
meaning and result type) and call hand-written sm_100a CUDA through a ctypes C ABI
(`include/pde_b200.h`, `libpde_b200.so`).  There is no CPU fallback."""
from .fields import TimeSeriesField, SolveResult, PlotResult  # noqa: F401
from . import _lib  # noqa: F401
from .solvers import (  # noqa: F401
    _solve_heat_1d_raw, _solve_heat_2d_raw, _solve_heat_3d_raw,
    _solve_elasticity_1d_static, _solve_elasticity_2d_static, _solve_elasticity_3d_static,
    _solve_heat_1d_cylindrical_raw, _solve_heat_1d_spherical_raw, _solve_heat_2d_cylindrical_raw,
    _solve_heat_2d_spherical_raw, _solve_heat_3d_spherical_raw,
    last_stats,
)
from . import mesh  # noqa: F401
from . import io  # noqa: F401

__all__ = ["TimeSeriesField", "SolveResult", "PlotResult", "mesh", "last_stats",
           "_solve_heat_1d_raw", "_solve_heat_2d_raw", "_solve_heat_3d_raw",
           "_solve_elasticity_1d_static", "_solve_elasticity_2d_static", "_solve_elasticity_3d_static"]
