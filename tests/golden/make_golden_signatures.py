"""Extract the tool / raw-solver signatures of the reference server into tests/golden/ (run in the
build container, where /root/reference exists; the JSON travels with the repo).

Only names, argument order and default values are recorded (via `ast`, the reference is not imported:
it needs FEniCS)."""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/fenics_mcp_server.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden",
                   "reference_signatures.json")


def sig(fn: ast.FunctionDef):
    args = fn.args.args
    defaults = [None] * (len(args) - len(fn.args.defaults)) + list(fn.args.defaults)
    out = []
    for a, d in zip(args, defaults):
        ent = {"name": a.arg, "annotation": ast.unparse(a.annotation) if a.annotation else None}
        if d is not None:
            ent["default"] = ast.literal_eval(d)
            ent["has_default"] = True
        else:
            ent["has_default"] = False
        out.append(ent)
    return out


tree = ast.parse(open(REF, encoding="utf-8").read())
tools, raw = {}, {}
server_name = None
for node in tree.body:
    if isinstance(node, ast.FunctionDef):
        is_tool = any(isinstance(d, ast.Call) and ast.unparse(d.func) == "mcp.tool" for d in node.decorator_list)
        if is_tool:
            tools[node.name] = {"args": sig(node), "returns": ast.unparse(node.returns) if node.returns else None}
        elif node.name.startswith("_solve_"):
            raw[node.name] = {"args": sig(node)}
    if isinstance(node, ast.Assign) and ast.unparse(node.targets[0]) == "mcp":
        server_name = ast.literal_eval(node.value.args[0])
json.dump({"source": "fenics_mcp_server.py (reference)", "server_name": server_name, "tools": tools, "raw": raw},
          open(OUT, "w"), indent=1, sort_keys=True)
print(OUT, len(tools), "tools", len(raw), "raw solvers")
