"""NumPy model of the elasticity GMG-PCG (test infrastructure / tuning aid; uses the oracle assembly).

Questions it answered in round 1 (cantilever 80x16x16, gravity, rtol 1e-10): point-Jacobi vs 3x3 block-Jacobi
Chebyshev smoothing (21 vs 20 PCG iterations: no gain), smoothing degree nu = 1..4 and Chebyshev ratios (nu = 2,
ratio 8-12 is the cheapest), scaling of the coarse-grid correction (1.0 is best: the re-discretised coarse
operators are the Galerkin ones), and hierarchy depth (two-grid with an exact coarse solve: 16 iterations, full
V-cycle: 21) - the iteration count is set by the P1 coarse spaces, not by the smoother."""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fem_oracle as fo
from oracle.gmg_model import prolongation, cheby_coefs

def build(n, L=(1.0, 0.2, 0.2), E=210e9, nu=0.3, min_cells=2):
    lam, mu = fo.lame(E, nu, 3)
    levels = []
    n = list(n)
    while True:
        m = fo.make_mesh(3, L, n)
        A = fo.assemble_elasticity(m, lam, mu).tocsr()      # interleaved 3*v + c
        nv = m.nv
        lv = type('L', (), {})()
        lv.n = list(n); lv.nv = nv
        clampv = fo.near(m.coords[:, 0], 0.0)
        lv.free = np.repeat(~clampv, 3)
        F = sp.diags(lv.free.astype(float))
        lv.Am = (F @ A @ F).tocsr()
        d = A.diagonal()
        lv.dinv_pt = np.where(lv.free, 1.0 / d, 0.0)
        # 3x3 diagonal blocks
        blk = np.zeros((nv, 3, 3))
        Ac = A.tocoo()
        same = (Ac.row // 3) == (Ac.col // 3)
        np.add.at(blk, (Ac.row[same] // 3, Ac.row[same] % 3, Ac.col[same] % 3), Ac.data[same])
        binv = np.linalg.inv(blk)
        binv[clampv] = 0.0
        lv.binv = binv
        lv.m = m
        levels.append(lv)
        can = all(k % 2 == 0 and k // 2 >= min_cells for k in n)
        if not can: break
        n = [k // 2 for k in n]
    for a in levels[:-1]:
        Ps = prolongation(a.n)
        a.P = sp.kron(Ps, sp.eye(3)).tocsr()
    lc = levels[-1]
    idx = np.nonzero(lc.free)[0]
    lc.idx = idx
    lc.dense = np.linalg.inv(lc.Am[idx][:, idx].toarray())
    return levels

def apply_dinv(lv, r, block):
    if not block: return lv.dinv_pt * r
    return np.einsum('vij,vj->vi', lv.binv, r.reshape(-1, 3)).ravel()

def est_lmax(lv, block, iters=12):
    x = lv.free * (np.sin(0.37 * np.arange(lv.Am.shape[0])) + 0.5)
    lam = 0
    for _ in range(iters):
        y = apply_dinv(lv, lv.Am @ x, block)
        lam = np.linalg.norm(y) / np.linalg.norm(x)
        x = y / np.linalg.norm(y)
    return 1.1 * lam

def smooth(lv, b, x, sweeps, ratio, block):
    d = np.zeros_like(b)
    for c1, c2 in cheby_coefs(lv.lmax, ratio, sweeps):
        r = lv.free * (b - lv.Am @ x)
        d = c1 * d + c2 * apply_dinv(lv, r, block)
        x = x + d
    return x

def vcycle(levels, b, nu, ratio, block, l=0):
    lv = levels[l]
    if l == len(levels) - 1:
        x = np.zeros_like(b); x[lv.idx] = lv.dense @ b[lv.idx]; return x
    x = smooth(lv, b, np.zeros_like(b), nu, ratio, block)
    r = lv.free * (b - lv.Am @ x)
    nxt = levels[l + 1]
    bc = nxt.free * (lv.P.T @ r)
    xc = vcycle(levels, bc, nu, ratio, block, l + 1)
    x = x + lv.free * (lv.P @ xc)
    return smooth(lv, b, x, nu, ratio, block)

def pcg(levels, b, block, nu=2, ratio=8.0, rtol=1e-10):
    lv = levels[0]
    b = lv.free * b
    x = np.zeros_like(b); r = b.copy()
    z = vcycle(levels, r, nu, ratio, block); p = z.copy(); rho = r @ z; bn = np.linalg.norm(b)
    for it in range(1, 200):
        q = lv.Am @ p; a = rho / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) / bn <= rtol: return it
        z = vcycle(levels, r, nu, ratio, block); rn = r @ z; p = z + (rn / rho) * p; rho = rn
    return -1

for n in ():
    levels = build(n)
    lv = levels[0]
    load = np.zeros(lv.Am.shape[0]); load[2::3] = -76518.0 * fo.lumped_load(lv.m)
    for block in (False, True):
        for l in levels: l.lmax = est_lmax(l, block)
        for ratio in (8.0, 4.0, 16.0):
            print(n, 'block' if block else 'point', 'ratio', ratio, 'lmax', round(levels[0].lmax, 3), 'iters', pcg(levels, load, block, ratio=ratio), flush=True)
print("--- nu sweep (point Jacobi)")
levels = build([80, 16, 16])
lv = levels[0]
load = np.zeros(lv.Am.shape[0]); load[2::3] = -76518.0 * fo.lumped_load(lv.m)
for l in levels: l.lmax = est_lmax(l, False)
for nu in (1, 2, 3, 4):
    for ratio in (4.0, 8.0, 12.0):
        it = pcg(levels, load, False, nu=nu, ratio=ratio)
        print('nu', nu, 'ratio', ratio, 'iters', it, 'cost', it * (1 + 1 + (2 * nu - 0.7) + 3), flush=True)
print("--- coarse correction scaling")
def vcycle_s(levels, b, nu, ratio, alpha, l=0):
    lv = levels[l]
    if l == len(levels) - 1:
        x = np.zeros_like(b); x[lv.idx] = lv.dense @ b[lv.idx]; return x
    x = smooth(lv, b, np.zeros_like(b), nu, ratio, False)
    r = lv.free * (b - lv.Am @ x)
    nxt = levels[l + 1]
    xc = vcycle_s(levels, nxt.free * (lv.P.T @ r), nu, ratio, alpha, l + 1)
    x = x + alpha * (lv.free * (lv.P @ xc))
    return smooth(lv, b, x, nu, ratio, False)
def pcg_s(levels, b, alpha, nu=2, ratio=8.0, rtol=1e-10):
    lv = levels[0]
    b = lv.free * b
    x = np.zeros_like(b); r = b.copy()
    z = vcycle_s(levels, r, nu, ratio, alpha); p = z.copy(); rho = r @ z; bn = np.linalg.norm(b)
    for it in range(1, 200):
        q = lv.Am @ p; a = rho / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) / bn <= rtol: return it
        z = vcycle_s(levels, r, nu, ratio, alpha); rn = r @ z; p = z + (rn / rho) * p; rho = rn
    return -1
for alpha in (1.0, 1.25, 1.5, 2.0, 3.0):
    print('alpha', alpha, 'iters', pcg_s(levels, load, alpha), flush=True)
# two-level check: exact coarse solve on level 1 -> is the hierarchy depth the limit?
print("--- depth check")
for depth in (2, 3, len(levels)):
    lv2 = levels[:depth]
    lc = lv2[-1]
    idx = np.nonzero(lc.free)[0]; lc.idx = idx
    lu = spl.splu(lc.Am[idx][:, idx].tocsc())
    class D:
        def __init__(s, lu): s.lu = lu
        def __matmul__(s, v): return s.lu.solve(v)
    lc.dense = D(lu)
    print('levels', depth, 'iters', pcg_s(lv2, load, 1.0), flush=True)
