#!/usr/bin/env python
"""Dispatcher-level entry without the LLM (`dispatcher_agent.py:97-1144`).

The reference turns a natural-language request into a `PDEParameters` object (LLM), and its `DispatcherAgent` then picks
the MCP tool and builds its kwargs with plain Python (`_build_*_args`).  This CLI feeds that second half directly: it
imports the reference's `dispatcher_agent.py` and `pde_schema.py` UNMODIFIED from a checkout (the LangChain / OpenAI
imports are stubbed - there is no network), hands the agent an in-process "MCP client" whose tools are this repository's
`fenics_mcp_server.py` functions, and runs `DispatcherAgent.dispatch(PDEParameters(**json))`: tool selection, argument
mapping, solve on the GPU, plot - the orchestrator's flow minus the language model.

    python dispatch_cli.py --reference /path/to/PDE-Solver --params '{"pde_type": "heat", "dim": 1, ...}'
    python dispatch_cli.py --reference /path/to/PDE-Solver --params-file case.json [--dry-run] [--time]

--dry-run prints the tool name and kwargs the reference dispatcher would call and stops (no GPU needed).
--time repeats nothing: it reports the solver statistics (`last_stats()`) of the call next to the wall time."""
import argparse
import asyncio
import importlib.util
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference_dispatcher(ref_root):
    """(DispatcherAgent, PDEParameters) from the reference checkout, with the network-bound imports stubbed."""
    class _Nothing:
        def __init__(self, *a, **k):
            pass
    for mod, attrs in (("langchain_openai", {"ChatOpenAI": _Nothing}),
                       ("langchain_core", {}), ("langchain_core.messages", {"HumanMessage": _Nothing, "SystemMessage": _Nothing}),
                       ("langchain_mcp_adapters", {}), ("langchain_mcp_adapters.client", {"MultiServerMCPClient": _Nothing})):
        try:
            importlib.import_module(mod)
        except Exception:
            _stub(mod, **attrs)
    out = {}
    for name in ("pde_schema", "dispatcher_agent"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod            # dispatcher_agent does `from pde_schema import PDEParameters`
        spec.loader.exec_module(mod)
        out[name] = mod
    return out["dispatcher_agent"].DispatcherAgent, out["pde_schema"].PDEParameters


class _Tool:
    def __init__(self, name, fn, record=None):
        self.name, self.fn, self.record = name, fn, record

    async def ainvoke(self, args):
        if self.record is not None:
            self.record.append((self.name, dict(args)))
            if self.fn is None:        # dry run: pretend to have solved
                return {"data_file": f"data/{self.name}_dryrun.pkl", "html_path": "plots/dryrun.html", "dim": 0, "meta": {}}
        res = self.fn(**{k: v for k, v in args.items()})
        return res if isinstance(res, dict) else dict(vars(res))


class InProcessClient:
    """Stands in for MultiServerMCPClient: the tools are plain function calls into fenics_mcp_server.py."""

    def __init__(self, dry_run=False):
        self.calls = []
        self.dry_run = dry_run

    async def get_tools(self):
        if self.dry_run:
            names = ["solve_heat_1D", "solve_heat_2D", "solve_heat_3D", "solve_heat_3D_spherical", "solve_heat_1D_cylindrical",
                     "solve_heat_1D_spherical", "solve_heat_2D_cylindrical", "solve_heat_2D_spherical",
                     "solve_elasticity_1D_static", "solve_elasticity_2D_static", "solve_elasticity_3D_static",
                     "plot_time_series_field_from_file"]
            return [_Tool(n, None, self.calls) for n in names]
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import fenics_mcp_server as srv
        tools = []
        for n in dir(srv):
            if n.startswith("solve_") or n.startswith("plot_time_series_field"):
                tools.append(_Tool(n, getattr(srv, n), self.calls))
        return tools


def dispatch(ref_root, params, dry_run=False):
    Agent, PDEParameters = load_reference_dispatcher(ref_root)
    client = InProcessClient(dry_run=dry_run)
    agent = Agent(mcp_client=client, llm=object())
    p = PDEParameters(**params)
    t0 = time.perf_counter()
    result = asyncio.run(agent.dispatch(p))
    return result, client.calls, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of ziyu0425/PDE-Solver (dispatcher_agent.py, pde_schema.py)")
    ap.add_argument("--params", help="PDEParameters fields as JSON")
    ap.add_argument("--params-file")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    params = json.loads(args.params) if args.params else json.load(open(args.params_file))
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # the reference prints debug lines: keep stdout for the one JSON answer
    result, calls, wall = dispatch(args.reference, params, args.dry_run)
    out = {"calls": [{"tool": n, "args": a} for n, a in calls], "wall_s": wall}
    if "error" in result:
        out["error"] = result["error"]
    else:
        out.update(data_file=result.get("data_file"), html_path=result.get("html_path"))
    if args.time and not args.dry_run:
        import pde_solver_b200 as P
        out["stats"] = P.last_stats()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out, default=str))
    return 1 if "error" in out else 0


if __name__ == "__main__":
    sys.exit(main())
