"""All GPUs of the box from ONE tool process (SURVEY §8e launch model).

The orchestrator starts a single `python fenics_mcp_server.py` (multi_agent_orchestrator.py:70-78), so the drop-in
cannot rely on torchrun.  With PDE_B200_GPUS=N (N > 1) the tool process keeps rank 0 and spawns N-1 worker processes,
one per GPU, that join its NCCL communicator (unique id over a pipe) and execute the SAME slab-partitioned C-ABI call
(pde_heat_solve / pde_elasticity_solve) on their z-slab.  Results come back through POSIX shared memory in the natural
(z-slowest) vertex order, so the caller sees exactly the single-GPU arrays.  Workers are started on first use and stay
alive for the life of the tool process.

Only the 3-D box tools use it (solve_heat_3D box branch, solve_elasticity_3D_static); everything else is small.
"""
import atexit
import ctypes as C
import os
import pickle
import struct
import subprocess
import sys
from multiprocessing import resource_tracker, shared_memory

import numpy as np

from . import _lib


def requested_gpus():
    try:
        return max(1, int(os.environ.get("PDE_B200_GPUS", "1")))
    except ValueError:
        return 1


def usable(n_cells_z, world=None):
    """Multi-GPU applies when asked for and every rank gets at least four cell layers.  With `world` given the answer
    is pure arithmetic (tests); otherwise PDE_B200_GPUS must not exceed the devices this process can see - a worker that
    cannot open its GPU would leave the others waiting in the communicator set-up."""
    if world is None:
        world = requested_gpus()
        if world > 1 and world > _lib.device_count():
            raise _lib.PdeError(f"PDE_B200_GPUS={world} but only {_lib.device_count()} CUDA device(s) are visible")
    return world > 1 and n_cells_z >= 4 * world


# ---------------------------------------------------------------------------------------------- wire format
def _send(f, obj):
    b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    f.write(struct.pack("<Q", len(b)))
    f.write(b)
    f.flush()


def _recv(f):
    h = f.read(8)
    if len(h) < 8:
        raise EOFError("multi-GPU worker pipe closed")
    (n,) = struct.unpack("<Q", h)
    return pickle.loads(f.read(n))


def _attach(name):
    shm = shared_memory.SharedMemory(name=name)
    try:   # the creator owns the segment: keep this process's resource tracker from unlinking it at exit
        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
    return shm


# ---------------------------------------------------------------------------------------------- the per-rank work
def _slab(n, rank, world):
    z0, nzl, nzg = _lib.slab_partition(3, n, rank, world)
    plane = (n[0] + 1) * (n[1] + 1)
    return z0 * plane, nzl * plane, nzg * plane


def _run_heat(ctx, rank, world, p_bytes, o_bytes, nsnap, out):
    """out: global [nsnap][nv] float64 array (shared memory); every rank fills its slab columns."""
    p = _lib.HeatParams.from_buffer_copy(p_bytes)
    o = _lib.SolverOpts.from_buffer_copy(o_bytes)
    off, nloc, _ = _slab(list(p.n), rank, world)
    values = np.empty((nsnap, nloc), dtype=np.float64)
    times = np.empty(nsnap, dtype=np.float64)
    st = _lib.Stats()
    _lib.check(_lib.lib().pde_heat_solve(ctx.handle, C.byref(p), C.byref(o), None, _lib.ptr(values), _lib.ptr(times),
                                         C.byref(st)))
    out[:, off:off + nloc] = values
    return st.as_dict(), times


def _run_elasticity(ctx, rank, world, p_bytes, o_bytes, want_disp, out, disp):
    p = _lib.ElastParams.from_buffer_copy(p_bytes)
    o = _lib.SolverOpts.from_buffer_copy(o_bytes)
    off, nloc, _ = _slab(list(p.n), rank, world)
    vm = np.empty(nloc, dtype=np.float64)
    dl = np.empty((nloc, 3), dtype=np.float64) if want_disp else None
    st, sp = _lib.Stats(), _lib.Stats()
    _lib.check(_lib.lib().pde_elasticity_solve(ctx.handle, C.byref(p), C.byref(o), _lib.ptr(vm), _lib.ptr(dl),
                                               C.byref(st), C.byref(sp)))
    out[off:off + nloc] = vm
    if want_disp:
        disp[off:off + nloc, :] = dl
    return st.as_dict(), sp.as_dict()


def _execute(ctx, rank, world, msg):
    kind = msg["kind"]
    segs = [_attach(nm) for nm in msg["shm"]]
    try:
        if kind == "heat":
            out = np.ndarray((msg["nsnap"], msg["nv"]), dtype=np.float64, buffer=segs[0].buf)
            return _run_heat(ctx, rank, world, msg["p"], msg["o"], msg["nsnap"], out)
        if kind == "elasticity":
            out = np.ndarray((msg["nv"],), dtype=np.float64, buffer=segs[0].buf)
            disp = np.ndarray((msg["nv"], 3), dtype=np.float64, buffer=segs[1].buf) if msg["want_disp"] else None
            return _run_elasticity(ctx, rank, world, msg["p"], msg["o"], msg["want_disp"], out, disp)
        raise ValueError(f"unknown command {kind}")
    finally:
        for s in segs:
            s.close()


def _worker_main(rank, world, device):
    fin, fout = sys.stdin.buffer, os.fdopen(os.dup(1), "wb")
    os.dup2(2, 1)                      # whatever a library prints goes to stderr, the pipe carries only replies
    hello = _recv(fin)
    try:                               # say whether this rank has its GPU BEFORE anyone enters the collective set-up
        ctx = _lib.Context(device)
    except Exception as e:  # noqa: BLE001
        _send(fout, ("err", f"rank {rank} (device {device}): {e}"))
        return
    _send(fout, ("ctx", rank))
    ctx.comm_init(rank, world, hello["uid"], hello["nccl"])
    _send(fout, ("ready", rank))
    while True:
        try:
            msg = _recv(fin)
        except EOFError:
            break
        if msg.get("kind") == "quit":
            break
        try:
            _send(fout, ("ok", _execute(ctx, rank, world, msg)))
        except Exception as e:  # noqa: BLE001 - reported to the tool process, which raises
            _send(fout, ("err", f"rank {rank}: {e}"))


# ---------------------------------------------------------------------------------------------- the pool (rank 0)
class Pool:
    def __init__(self, world):
        self.world = world
        self.broken = False
        path = _lib.nccl_library_path()
        uid = _lib.nccl_unique_id(path)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env.pop("PDE_B200_GPUS", None)
        self.procs = []
        for r in range(1, world):
            pr = subprocess.Popen([sys.executable, "-m", "pde_solver_b200.multi", "--worker", str(r), str(world), str(r)],
                                  stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env, cwd=root)
            _send(pr.stdin, {"uid": uid, "nccl": path})
            self.procs.append(pr)
        try:
            self.ctx = _lib.Context(0)
            for pr in self.procs:                        # every rank holds a device context before the collective
                self._expect(pr, "ctx")
            self.ctx.comm_init(0, world, uid, path)      # collective with the workers' comm_init
            for pr in self.procs:
                self._expect(pr, "ready")
        except Exception:
            self.close(kill=True)
            raise
        atexit.register(self.close)

    @staticmethod
    def _expect(pr, want):
        try:
            tag, payload = _recv(pr.stdout)
        except EOFError:
            raise _lib.PdeError(f"multi-GPU worker exited during start-up (exit code {pr.poll()})") from None
        if tag != want:
            raise _lib.PdeError(f"multi-GPU worker failed to start: {payload}")

    def run(self, msg, shapes):
        """Allocate the shared result arrays, run the command on every rank, return (rank-0 result, arrays)."""
        segs, arrs = [], []
        try:
            for shp in shapes:
                n = int(np.prod(shp)) * 8
                s = shared_memory.SharedMemory(create=True, size=max(n, 8))
                segs.append(s)
                arrs.append(np.ndarray(shp, dtype=np.float64, buffer=s.buf))
            msg = dict(msg, shm=[s.name for s in segs])
            for pr in self.procs:
                _send(pr.stdin, msg)
            err = None
            try:
                mine = _execute(self.ctx, 0, self.world, msg)
            except Exception as e:  # noqa: BLE001
                err, mine = str(e), None
            for pr in self.procs:
                try:
                    tag, payload = _recv(pr.stdout)
                except EOFError:                       # the rank is gone: this pool cannot serve another call
                    tag, payload = "err", f"worker exited (exit code {pr.poll()})"
                    self.broken = True
                if tag != "ok":
                    err = err or payload
            if err:
                raise _lib.PdeError(f"multi-GPU solve failed: {err}")
            return mine, [np.array(a) for a in arrs]          # private copies: the segments go away below
        finally:
            for s in segs:
                s.close()
                s.unlink()

    def close(self, kill=False):
        for pr in self.procs:
            try:
                _send(pr.stdin, {"kind": "quit"})
                pr.stdin.close()
            except Exception:
                pass
        for pr in self.procs:
            try:
                if kill:                               # a rank may be stuck in the communicator set-up: do not wait
                    pr.kill()
                pr.wait(timeout=10)
            except Exception:
                pr.kill()
        self.procs = []


_pool = None


def pool():
    global _pool
    world = requested_gpus()
    if _pool is None or _pool.world != world or _pool.broken:
        if _pool is not None:
            _pool.close(kill=_pool.broken)
        _pool = Pool(world)
    return _pool


def heat_solve(p, o, nsnap, nv):
    (st, times), (values,) = pool().run({"kind": "heat", "p": bytes(p), "o": bytes(o), "nsnap": nsnap, "nv": nv},
                                        [(nsnap, nv)])
    st["gpus"] = requested_gpus()
    return st, values, times


def elasticity_solve(p, o, nv, want_disp):
    shapes = [(nv,), (nv, 3)] if want_disp else [(nv,)]
    (st, sp), arrs = pool().run({"kind": "elasticity", "p": bytes(p), "o": bytes(o), "nv": nv, "want_disp": want_disp},
                                shapes)
    st["gpus"] = requested_gpus()
    return st, sp, arrs[0], (arrs[1] if want_disp else None)


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "--worker":
        _worker_main(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
