mkdir -p gpurun_out
for rep in 1 2; do for lib in ab/libpde_base.so ab/libpde_e640.so; do
  echo "== $lib"; PDE_B200_LIB=$lib timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,1,3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('  ', d['mode'], d['ms'], d['GBps'])
    except Exception: print(l.strip()[:200])"
  PDE_B200_LIB=$lib timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('   solve', d['iters'], round(d['solve_ms'],1))
    except Exception: print(l.strip()[:200])"
done; done
