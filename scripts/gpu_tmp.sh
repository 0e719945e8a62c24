mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "manufactured or heat_3d or smoke or superposition" 2>&1 | tail -3 ) ; 
for rep in 1 2; do for lib in ab/libpde_base.so ab/libpde_new.so; do
  PDE_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu --no-elasticity > gpurun_out/ab.json 2>gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read()); print('$lib', 'heat ms/step', round(d['ms_per_step'],2))"
done; done
