mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
echo "rc=$?"; tail -5 gpurun_out/r02_bench_n1_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_a.json').read())
for k in ('value','ms_per_step','cg_iters_per_step','true_relres','roofline_step','e2e'): print(k, d.get(k))
print('roofline', d['roofline'])
print('sweeps', d['sweeps'])
e=d['elasticity']; print('elast', {k:e[k] for k in e if k not in ('sweeps',)}); print(e['sweeps'])
print('configs', d['configs'])
PY
