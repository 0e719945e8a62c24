mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -k "elasticity or smoke or sweep_modes or operator_apply or manufactured" 2>&1 | tail -6 ) > gpurun_out/r02_t21.log 2>&1; cat gpurun_out/r02_t21.log
for v in 1 0; do
echo "FACE_TMA=$v"; PDE_B200_FACE_TMA=$v timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('  ', d['mode'], d['ms'], d['GBps'])
    except Exception: print(l.strip()[:200])"
PDE_B200_FACE_TMA=$v timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('  solve', d['iters'], round(d['solve_ms'],1), d['relres'])
    except Exception: print(l.strip()[:200])"
done
