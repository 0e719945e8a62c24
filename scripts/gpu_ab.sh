# A/B of library builds within ONE call (boxes differ by a few per cent): scripts/gpu_ab.sh libA.so libB.so ...
mkdir -p gpurun_out
for rep in 1 2; do for lib in "$@"; do
  PDE_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/ab.json 2>gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read()); e=d['elasticity']; print('$lib', 'heat ms/step', round(d['ms_per_step'],2), 'elast solve ms', round(e['solve_ms'],1), 'apply frac', round(e['roofline']['frac'],3))"
done; done
