"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mgpu_check.py

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
if rank == 0:
    print("MGPU OK" if ok else "MGPU FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
