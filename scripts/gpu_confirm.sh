mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final.json').read())
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],'step',d['roofline_step']['frac'],d['roofline_step']['frac_by_round1_accounting_183B'])
e=d['elasticity']; print('elast solve',e['solve_ms'],'apply',e['roofline']['frac'],'step',e['roofline_step']['frac'],e['roofline_step']['frac_by_round1_accounting_247B'],'proj',e['projection_ms'])
print('clocks',d['clocks'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
PY
