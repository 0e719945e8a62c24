mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -k "elasticity or manufactured or smoke or sweep_modes" 2>&1 | tail -6 ) > gpurun_out/r02_t18.log 2>&1; cat gpurun_out/r02_t18.log
for v in 1 0; do
PDE_B200_E_FIRST2=$v timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('FIRST2=$v', d['iters'], round(d['solve_ms'],1), d['relres'])
    except Exception: print(l.strip()[:200])"
done
