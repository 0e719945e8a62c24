#!/usr/bin/env python
"""Headline benchmark of the hot path (BASELINE.json): 3D P1 heat, backward Euler, config 4.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

One "step" = one backward-Euler time step (a full linear solve to rtol 1e-10) of the heat equation
on the unit-cube-per-GPU 512^3 P1 mesh (135 M dofs per GPU; weak scaling stacks slabs along z).
`value` = dofs advanced one time step per second (GDOF/s), state resident in HBM, timed with CUDA
events on the library's stream, max over ranks.  Extra keys report CG iterations/s, the operator
micro-benchmark and its HBM roofline fraction, the end-to-end number through the C ABI with host
buffers, a bounded CPU baseline and the clocks seen during the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells per GPU, L per GPU block, kappa, dt, T_initial, T_boundary)
    "heat3d_512": dict(n=(512, 512, 512), kappa=1.0, dt=0.01, T_initial=20.0, T_boundary=0.0),
    "heat3d_256": dict(n=(256, 256, 256), kappa=1.0, dt=0.01, T_initial=20.0, T_boundary=0.0),
    "heat3d_128": dict(n=(128, 128, 128), kappa=1.0, dt=0.01, T_initial=20.0, T_boundary=0.0),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(n, steps, warmup):
    from oracle.reference_arm import HeatReference3D
    ref = HeatReference3D(n)
    sec = ref.run(steps, warmup)
    return ref.ndofs, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 32 if (args.steps + args.warmup) <= 8 else 24
    ndofs, sec = cpu_reference(n, args.steps, args.warmup)
    val = ndofs * args.steps / sec / 1e9
    out = {
        "impl": "reference", "metric": "GDOF/s", "value": val, "unit": "GDOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": f"{n}^3 cells ({ndofs} dofs), same kappa/dt/IC/BC"},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": 1, "kind": "port",
                         "sample": f"3D heat {n}^3 cells, {args.steps} backward-Euler steps, per step: assemble A and b, "
                                   "row-wise Dirichlet, SuperLU factorise+solve (what DOLFIN solve() does each step); "
                                   "SciPy restatement, single-threaded"},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    args.emit_restore()
    print(json.dumps(out), flush=True)
    return 0


def run_native(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    import numpy as np
    from pde_solver_b200 import _lib

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    if world > 1:
        import torch
        path = _lib.nccl_library_path()
        uid = [_lib.nccl_unique_id(path) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0], path)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w = WORKLOADS[args.workload]
    n = [w["n"][0], w["n"][1], w["n"][2] * world]      # weak scaling: stack per-GPU blocks along z
    L = [1.0, 1.0, 1.0 * world]
    bc = _lib.make_bc({f: w["T_boundary"] for f in range(6)})
    precond = args.precond
    if world > 1 and precond != "jacobi" and not args.force_precond:
        precond = args.precond
    opts = _lib.make_opts(rtol=args.rtol, precond=precond)
    hs = _lib.HeatStepper(ctx, 3, n, L, w["kappa"], w["dt"], T_initial=w["T_initial"], bc=bc, opts=opts)
    ndofs = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)

    # ---- warm-up, then the timed region (inputs resident in HBM) ----
    for _ in range(args.warmup):
        hs.step(1)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    st = hs.step(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(st["solve_ms"])
    value = ndofs * args.steps / (ms / 1e3) / 1e9
    iters = st["iters_total"]

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + step + D2H per step) ----
    # every step: H2D copy of that step's input field from pinned host memory, one backward-Euler solve, D2H read of
    # the result into pinned host memory.  (1) pipelined: pde_heat_advance_batch over e_steps independent requests;
    # the upload of request k+1 and the download of result k-1 run on copy streams while request k is solved.
    # (2) serial, for comparison: set_state / step / get_state, the result being the next step's input.
    hbuf = _lib.PinnedArray(hs.nloc)
    hs.get_state(hbuf.array)
    e_steps = max(1, args.steps)
    h_in = [_lib.PinnedArray(hs.nloc) for _ in range(2)]
    h_out = [_lib.PinnedArray(hs.nloc) for _ in range(2)]
    for b in h_in:
        b.array[:] = hbuf.array
    ins = [h_in[k % 2].array for k in range(e_steps)]
    outs = [h_out[k % 2].array for k in range(e_steps)]
    hs.advance_batch(ins[:2], outs[:2])                 # untimed: allocates the staging buffers and streams
    barrier()
    t0 = time.perf_counter()
    hs.advance_batch(ins, outs)
    barrier()
    e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e = ndofs * e_steps / e_sec / 1e9
    s_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(s_steps):
        hs.set_state(hbuf.array)
        hs.step(1)
        hs.get_state(hbuf.array)
    barrier()
    s_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_serial = ndofs * s_steps / s_sec / 1e9
    bytes_dir = int(sum_over_ranks(hs.nloc * 8))
    for b in h_in + h_out:
        b.free()
    hs.close()
    hbuf.free()

    # ---- dominant kernel: matrix-free operator apply y = (M + dt k K) x, 16 B/dof algorithmic ----
    peaks, peak_kind = measured_peaks()
    op = _lib.op_params("heat", 3, n, L, alpha=1.0, beta=w["dt"] * w["kappa"], bc=bc)
    barrier()
    op_ms, op_nd = _lib.op_bench(ctx, op, reps=20, warmup=3)
    op_ms = max_over_ranks(op_ms)
    ach = 16.0 * op_nd / (op_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "kernel": "heat operator apply (+fused dot)",
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "ms_per_launch": op_ms,
                "per_gpu_dofs": op_nd, "nominal_8TBs_frac": ach / 8000.0}

    # ---- halo exchange (N > 1): one 513x513 plane each way per z-neighbour over NVLink ----
    halo = None
    if world > 1:
        barrier()
        h_ms, h_bytes = _lib.halo_bench(ctx, 3, n, 1, reps=50)
        h_ms = max_over_ranks(h_ms)
        plane_bytes = (((n[0] + 2 + 3) // 4) * 4) * (n[1] + 2) * 8      # one padded vertex plane
        gbs = plane_bytes / (h_ms / 1e3) / 1e9 if h_ms > 0 else None
        halo = {"ms_per_exchange": h_ms, "plane_bytes": plane_bytes, "bytes_sent_per_rank_max": int(max_over_ranks(h_bytes)),
                "gbs_per_direction": gbs, "nvlink_peer_copy_peak_gbs": 770.0,
                "frac_of_nvlink": gbs / 770.0 if gbs else None,
                "path": "nccl send/recv" if os.environ.get("PDE_B200_HALO") == "nccl" else
                        "peer-memory mailbox kernel over NVLink (cudaIpc), NCCL send/recv as fallback",
                "note": "one contiguous plane per neighbour; latency-bound at this size (flag round trips, not bytes)"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- the other half of the headline metric: 3D elasticity, BASELINE config 5 (N=1, public host API) ----
    elast = None
    if world == 1 and not args.no_elasticity:
        import pde_solver_b200 as P
        en = (1280, 256, 256) if args.workload == "heat3d_512" else (320, 64, 64)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            P._solve_elasticity_3d_static(1.0, 0.2, 0.2, en[0], en[1], en[2], 210e9, 0.3, 0.0, 0.0, -76518.0, "stress",
                                          rtol=args.rtol, precond="gmg", as_arrays=True)
            wall = time.perf_counter() - t0
            es = P.last_stats()
            if best is None or es["solve_ms"] < best[0]["solve_ms"]:
                best = (es, wall)
        es, wall = best
        elast = {"workload": f"elast3d_{en[0]}x{en[1]}x{en[2]} cantilever, gravity, GMG-PCG rtol {args.rtol:g}",
                 "dofs": es["ndofs"], "cg_iters": es["iters_total"], "solve_ms": es["solve_ms"],
                 "setup_ms": es["setup_ms"], "projection_ms": es["projection"]["solve_ms"],
                 "gdofs_per_s": es["ndofs"] / (es["solve_ms"] / 1e3) / 1e9,
                 "cg_iters_per_s": es["iters_total"] / (es["solve_ms"] / 1e3),
                 "gdof_iters_per_s": es["ndofs"] * es["iters_total"] / (es["solve_ms"] / 1e3) / 1e9,
                 "converged": bool(es["converged"]), "final_relres": es["final_relres"],
                 "e2e_wall_s": wall, "levels": es["levels"]}

    # ---- bounded CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cn, csteps = 32, 3
        cnd, csec = cpu_reference(cn, csteps, 0)
        cpu = {"value": cnd * csteps / csec / 1e9, "unit": "GDOF/s", "cores": 1, "kind": "port",
               "sample": f"3D heat {cn}^3 cells ({cnd} dofs), {csteps} steps of assemble + row-wise BC + SuperLU "
                         "factorise/solve per step (SciPy restatement of the FEniCS path, single-threaded)"}

    out = {
        "metric": "GDOF/s", "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "cells": n, "dofs": ndofs, "dt": w["dt"], "kappa": w["kappa"],
                   "rtol": args.rtol, "precond": "gmg" if st["levels"] > 1 else "jacobi", "mg_levels": st["levels"],
                   "partition": f"z-slabs x{world}", "l2": "vectors (>=1 GiB each) exceed the 126 MB L2"},
        "cg_iters": iters, "cg_iters_per_s": iters / (ms / 1e3), "cg_iters_per_step": iters / args.steps,
        "converged": bool(st["converged"]), "final_relres": st["final_relres"],
        "operator_gdofs": op_nd * world / (op_ms / 1e3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": "GDOF/s", "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
                "steps": e_steps, "api": "pde_heat_advance_batch (3-stream pipeline, pinned host buffers)",
                "serial_value": e2e_serial, "serial_steps": s_steps},
        "gpu_launches": int(st["launches"]), "clocks": clocks, "elasticity": elast, "halo": halo,
    }
    args.emit_restore()
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)          # anything printed during teardown goes to stderr again
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly ONE JSON line: libraries that chat on stdout (NCCL prints its version / debug lines there)
    # are sent to stderr for the duration of the run, the result goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit_restore():
        sys.stdout.flush()
        os.dup2(real_stdout, 1)

    rc = 1
    try:
        rc = _main(emit_restore)
    finally:
        sys.stdout.flush()
    return rc


def _main(emit_restore):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="heat3d_512", choices=sorted(WORKLOADS))
    ap.add_argument("--precond", default="auto", choices=["auto", "gmg", "jacobi"])
    ap.add_argument("--force-precond", action="store_true")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-elasticity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    args.emit_restore = emit_restore
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
