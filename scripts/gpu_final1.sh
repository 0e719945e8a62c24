mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final.json').read())
print('value',d['value'],'ms/step',d['ms_per_step'],'iters/step',d['cg_iters_per_step'],'true',d['true_relres'])
print('roofline',d['roofline']['frac'],'step',d['roofline_step'])
print('e2e',d['e2e']['value'],'clocks',d['clocks'])
e=d['elasticity']; print('elast solve',e['solve_ms'],'iters',e['cg_iters'],'apply frac',e['roofline']['frac'],'proj',e['projection_ms'],'step',e['roofline_step']['frac'],'true',e['true_relres'])
print([ (s['kernel'],round(s['frac'],3)) for s in d['sweeps']]); print([ (s['kernel'],round(s['frac'],3)) for s in e['sweeps']])
print('configs',d['configs'])
PY
