mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -k "elasticity or smoke or scaling_laws" 2>&1 | tail -6 ) > gpurun_out/r02_t20.log 2>&1; cat gpurun_out/r02_t20.log
for v in 1 0; do
PDE_B200_CELL_2PASS=$v timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('2PASS=$v', d['iters'], round(d['solve_ms'],1), 'proj', round(d['proj_ms'],1), d['proj_iters'], 'vm_max', d['vm_max'])
    except Exception: print(l.strip()[:200])"
done
