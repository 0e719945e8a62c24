"""Dispatcher-level CLI (row n4): the reference's UNMODIFIED DispatcherAgent (dispatcher_agent.py:97-1144) picks the tool
and builds its kwargs from a PDEParameters object; every kwarg it passes must be accepted by this repository's tool of
that name.  Needs the reference checkout (skipped on the GPU box, where it does not exist)."""
import inspect
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "dispatcher_agent.py")), reason="no reference checkout")
sys.path.insert(0, ROOT)

CASES = [
    # BASELINE config 1: the dispatcher derives dt = 0.01 and 200 steps for L = 2, kappa = 1 (dispatcher_agent.py:393-404)
    (dict(pde_type="heat", dim=1, domain_size={"length": 2.0}, nx=100, bc_values={"T_left": 20.0, "T_right": 0.0},
          initial_value=0.0), "solve_heat_1D", dict(length=2.0, nx=100, T_left=20.0, T_right=0.0, dt=0.01, num_steps=200)),
    (dict(pde_type="heat", dim=2, domain_size={"Lx": 1.0, "Ly": 1.0}, nx=64, ny=64, bc_values={"T_boundary": 0.0},
          initial_value=20.0, dt=0.01, num_steps=100), "solve_heat_2D", dict(Lx=1.0, Ly=1.0, nx=64, ny=64, T_initial=20.0)),
    (dict(pde_type="heat", dim=3, domain_size={"Lx": 1.0, "Ly": 1.0, "Lz": 1.0}, nx=16, ny=16, nz=16,
          bc_values={"T_boundary": 0.0}, initial_value=20.0, dt=0.01, num_steps=50), "solve_heat_3D",
     dict(nx=16, ny=16, nz=16, T_boundary=0.0, T_initial=20.0, num_steps=50)),
    (dict(pde_type="elasticity", dim=3, domain_size={"Lx": 1.0, "Ly": 0.2, "Lz": 0.2}, nx=40, ny=8, nz=8,
          young_modulus=210e9, poisson_ratio=0.3), "solve_elasticity_3D_static", dict(Lx=1.0, Ly=0.2, Lz=0.2, E=210e9, nu=0.3)),
    (dict(pde_type="heat", dim=1, geometry_type="cylinder", domain_size={"r1": 0.1, "r2": 1.0}, nx=50,
          bc_values={"T_inner": 100.0, "T_outer": 20.0}, initial_value=20.0), "solve_heat_1D_cylindrical", {}),
]


@pytest.mark.parametrize("params,tool,expect", CASES)
def test_reference_dispatcher_drives_our_tools(params, tool, expect):
    import dispatch_cli
    import fenics_mcp_server as srv
    result, calls, _ = dispatch_cli.dispatch(REF, params, dry_run=True)
    assert "error" not in result, result
    assert calls[0][0] == tool
    for k, v in expect.items():
        assert calls[0][1][k] == pytest.approx(v), (k, calls[0][1])
    accepted = set(inspect.signature(getattr(srv, tool)).parameters)
    assert set(calls[0][1]) <= accepted, set(calls[0][1]) - accepted            # drop-in: no kwarg our tool rejects
    assert calls[1][0] == "plot_time_series_field_from_file"
    assert set(calls[1][1]) <= set(inspect.signature(srv.plot_time_series_field_from_file).parameters)
