"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py

Slab-partitioned 3D heat (GMG-PCG and Jacobi-PCG) against the CPU oracle (<= 1e-8 rel-L2), plus an
operator application against the oracle matrix.  Rank 0 prints 'MGPU OK'."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pde_solver_b200 import _lib  # noqa: E402
from oracle import fem_oracle as fo  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _lib.Context(local)
path = _lib.nccl_library_path()
uid = [_lib.nccl_unique_id(path) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0], path)


def gather(local_arr):
    """Concatenate the ranks' slabs (rank order = plane order)."""
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_arr.size], dtype=torch.int64, device="cuda"))
    sizes = [int(s.item()) for s in sizes]
    bufs = [torch.zeros(max(sizes), dtype=torch.float64, device="cuda") for _ in range(world)]
    mine = torch.zeros(max(sizes), dtype=torch.float64, device="cuda")
    mine[:local_arr.size] = torch.from_numpy(local_arr).cuda()
    dist.all_gather(bufs, mine)
    return np.concatenate([b[:s].cpu().numpy() for b, s in zip(bufs, sizes)])


ok = True
n = [16, 16, 8 * world * 2]
L = [1.0, 1.0, 1.0 * world]
steps = 3
ref = fo.solve_heat(3, L, n, 1.0, T_initial=20.0, dt=0.01, num_steps=steps, T_boundary=0.0) if rank == 0 else None
for precond in ("gmg", "jacobi"):
    bc = _lib.make_bc({f: 0.0 for f in range(6)})
    hs = _lib.HeatStepper(ctx, 3, n, L, 1.0, 0.01, T_initial=20.0, bc=bc, opts=_lib.make_opts(rtol=1e-10, precond=precond))
    st = hs.step(steps)
    u = np.empty(hs.nloc)
    hs.get_state(u)
    hs.close()
    full = gather(u)
    if rank == 0:
        err = fo.rel_l2(full, ref.values[-1])
        print(f"[mgpu] heat {n} x{world} {precond}: rel-L2 {err:.2e} iters {st['iters_total']} levels {st['levels']} "
              f"converged {st['converged']}", flush=True)
        ok = ok and err <= 1e-8 and st["converged"] == 1 and (precond == "jacobi" or st["levels"] > 1)
# ---- 2-D heat, slabs along y (internal z): wide enough for the register-marching kernel k_sweep2d ----
n2, L2 = [192, 48 * world], [1.0, 0.5 * world]
ref2 = fo.solve_heat(2, L2, n2, 0.8, T_initial=6.0, dt=0.02, num_steps=3, T_boundary=1.0, source_type="constant",
                     source_value=3.0) if rank == 0 else None
for precond in ("gmg", "jacobi"):
    bc2 = _lib.make_bc({f: 1.0 for f in range(4)})
    hs = _lib.HeatStepper(ctx, 2, n2, L2, 0.8, 0.02, T_initial=6.0, bc=bc2, source_value=3.0,
                          opts=_lib.make_opts(rtol=1e-10, precond=precond))
    st = hs.step(3)
    u = np.empty(hs.nloc)
    hs.get_state(u)
    hs.close()
    full = gather(u)
    if rank == 0:
        err = fo.rel_l2(full, ref2.values[-1])
        print(f"[mgpu] heat 2-D {n2} x{world} {precond}: rel-L2 {err:.2e} iters {st['iters_total']} levels {st['levels']}",
              flush=True)
        ok = ok and err <= 1e-8 and st["converged"] == 1
# ---- elasticity: cantilever under gravity, slab-partitioned GMG-PCG + von Mises projection (C ABI, local slabs) ----
import ctypes as C
en, eL = [16, 4, 4 * world * 2], [1.0, 0.25, 0.5 * world]
z0, nzl, nzg = _lib.slab_partition(3, en, rank, world)
nloc = (en[0] + 1) * (en[1] + 1) * nzl
ep = _lib.ElastParams()
ep.dim = 3
ep.n = _lib.i3(en)
ep.L = _lib.d3(eL)
ep.E, ep.nu = 210e9, 0.3
ep.body = _lib.d3([0.0, 0.0, -76518.0], 0.0)
ep.quantity, ep.plane_stress, ep.area = 0, 0, 1.0
vm = np.empty(nloc)
disp = np.empty((nloc, 3))
st, sp = _lib.Stats(), _lib.Stats()
o = _lib.make_opts(rtol=1e-10, precond="gmg")
_lib.check(_lib.lib().pde_elasticity_solve(ctx.handle, C.byref(ep), C.byref(o), _lib.ptr(vm), _lib.ptr(disp),
                                           C.byref(st), C.byref(sp)))
vm_full = gather(vm)
u_full = gather(disp.ravel())
if rank == 0:
    eref = fo.solve_elasticity(3, eL, en, 210e9, 0.3, body=[0, 0, -76518.0], quantity="stress")
    e_vm = fo.rel_l2(vm_full, eref.values[0])
    e_u = fo.rel_l2(u_full.reshape(-1, 3), eref.aux["u"])
    print(f"[mgpu] elasticity {en} x{world} gmg: von Mises rel-L2 {e_vm:.2e}, displacement rel-L2 {e_u:.2e}, "
          f"iters {st.iters_total} levels {st.levels}", flush=True)
    ok = ok and e_vm <= 1e-8 and e_u <= 1e-8 and st.converged == 1
# ---- sizes that take the TMA sweep kernels and the temporally blocked smoother (2 halo planes): manufactured solutions
for kind, mn, mL, faces in (("heat", [96, 40, 32 * world], [1.0, 0.5, 0.4 * world], {f: 0.0 for f in range(6)}),
                            ("elasticity", [64, 16, 16 * world], [1.0, 0.25, 0.25 * world], {0: 0.0})):
    mp = _lib.op_params(kind, 3, mn, mL, 1.0, 0.01, 121.15e9, 80.77e9, bc=_lib.make_bc(faces))
    err, mst = _lib.op_manufactured(ctx, mp, _lib.make_opts(rtol=1e-11, precond="gmg"))
    if rank == 0:
        print(f"[mgpu] manufactured {kind} {mn} x{world}: rel-L2 {err:.2e} iters {mst['iters_total']} levels {mst['levels']} "
              f"true relres {mst['true_relres']:.1e}", flush=True)
        ok = ok and err <= 1e-8 and mst["converged"] == 1 and mst["levels"] > 1
# ---- cubic grids: the coarse levels end up with fewer planes than ranks (replicated levels below the slab levels) ----
for cn in ([64, 64, 64], [96, 96, 96]):
    if cn[2] % world:
        continue
    hs = _lib.HeatStepper(ctx, 3, cn, [1.0, 1.0, 1.0], 1.0, 0.01, T_initial=20.0, bc=_lib.make_bc({f: 0.0 for f in range(6)}),
                          opts=_lib.make_opts(rtol=1e-10, precond="gmg"))
    st = hs.step(2)
    u = np.empty(hs.nloc)
    hs.get_state(u)
    hs.close()
    full = gather(u)
    if rank == 0:
        # the N-GPU result against the 1-GPU result of the same global problem
        c1 = _lib.Context(local)
        h1 = _lib.HeatStepper(c1, 3, cn, [1.0, 1.0, 1.0], 1.0, 0.01, T_initial=20.0, bc=_lib.make_bc({f: 0.0 for f in range(6)}),
                              opts=_lib.make_opts(rtol=1e-10, precond="gmg"))
        s1 = h1.step(2)
        u1 = np.empty(h1.nloc)
        h1.get_state(u1)
        h1.close()
        d = fo.rel_l2(full, u1)
        print(f"[mgpu] heat {cn} x{world} vs 1 GPU: rel-L2 {d:.2e} (iters {st['iters_total']} vs {s1['iters_total']}, "
              f"levels {st['levels']}, true relres {st['true_relres']:.1e})", flush=True)
        ok = ok and d <= 1e-9 and st["converged"] == 1 and st["true_relres"] <= 1e-9
    dist.barrier()   # the other ranks must not start spinning on rank 0's halo flags while it solves alone
# elasticity, N GPUs against 1 GPU on the same global cantilever (sizes that take k_elast3d and the face-row kernel)
en2, eL2 = [64, 16, 16 * world], [1.0, 0.25, 0.25 * world]


def elast_solve(cx, rk, wd):
    z0_, nzl_, _ = _lib.slab_partition(3, en2, rk, wd)
    nl = (en2[0] + 1) * (en2[1] + 1) * nzl_
    q = _lib.ElastParams()
    q.dim = 3
    q.n = _lib.i3(en2)
    q.L = _lib.d3(eL2)
    q.E, q.nu = 210e9, 0.3
    q.body = _lib.d3([0.0, 0.0, -76518.0], 0.0)
    q.quantity, q.plane_stress, q.area = 0, 0, 1.0
    v, dd = np.empty(nl), np.empty((nl, 3))
    s_, p_ = _lib.Stats(), _lib.Stats()
    _lib.check(_lib.lib().pde_elasticity_solve(cx.handle, C.byref(q), C.byref(_lib.make_opts(rtol=1e-10, precond="gmg")),
                                               _lib.ptr(v), _lib.ptr(dd), C.byref(s_), C.byref(p_)))
    return v, dd, s_


vmN, dN, sN = elast_solve(ctx, rank, world)
vmN, dN = gather(vmN), gather(dN.ravel())
if rank == 0:
    c1 = _lib.Context(local)
    vm1, d1, s1 = elast_solve(c1, 0, 1)
    e1, e2 = fo.rel_l2(vmN, vm1), fo.rel_l2(dN, d1.ravel())
    print(f"[mgpu] elasticity {en2} x{world} vs 1 GPU: von Mises rel-L2 {e1:.2e}, displacement rel-L2 {e2:.2e} "
          f"(iters {sN.iters_total} vs {s1.iters_total}, true relres {sN.true_relres:.1e})", flush=True)
    ok = ok and e1 <= 1e-8 and e2 <= 1e-8 and sN.converged == 1
dist.barrier()
# ---- halo exchange self-check: sizes grow (the mailboxes are re-mapped collectively), 1 / 2 planes, 1 / 3 components
for n_h in ([40, 24, 16 * world], [96, 80, 24 * world], [512, 512, 8 * world]):
    for ncomp_h, depth_h in ((1, 1), (1, 2), (3, 1), (3, 2)):
        bad = _lib.halo_check(ctx, 3, n_h, ncomp_h, depth_h, reps=5)
        flags = [None] * world
        dist.all_gather_object(flags, bad)
        assert sum(flags) == 0, f"halo exchange mismatches {flags} for n={n_h} ncomp={ncomp_h} depth={depth_h}"
if rank == 0:
    print(f"[mgpu] halo self-check ok ({os.environ.get('PDE_B200_HALO', 'peer-memory')} path)", flush=True)
h_ms, h_bytes = _lib.halo_bench(ctx, 3, [512, 512, 64 * world], 1, reps=50)
if rank == 0:
    print(f"[mgpu] halo exchange 513x513 plane: {h_ms * 1e3:.1f} us, {516 * 514 * 8 / (h_ms / 1e3) / 1e9:.0f} GB/s per direction",
          flush=True)
    print("MGPU OK" if ok else "MGPU FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
