mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "heat_3d or manufactured or superposition or sine_mode or steady_heat or advance_batch or smoke" 2>&1 | tail -4 ) > gpurun_out/r02_t16.log 2>&1; cat gpurun_out/r02_t16.log
for v in 0 1; do
  PDE_B200_NO_RR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-elasticity --no-configs --no-cpu > gpurun_out/r02_bench_rr_$v.json 2>gpurun_out/r02_bench_rr_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_rr_$v.json').read()); print('NO_RR=$v', d['ms_per_step'], d['cg_iters_per_step'], d['roofline_step']['frac'], d['true_relres'])"
done
