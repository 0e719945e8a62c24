mkdir -p gpurun_out
N=${1:-8}; TAG=${2:-a}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n${N}_$TAG.json 2> gpurun_out/r02_bench_n${N}_$TAG.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n${N}_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}_$TAG.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','cg_iters_per_step','ms_per_iter','true_relres','halo'): print(k, d.get(k))
e=d['elasticity']; print('elast', {k:e[k] for k in ('cg_iters','solve_ms','ms_per_iter','halo_exchanges_per_iter','allreduces_per_iter')})
for k,v in d['strong'].items(): print('strong',k,{q:v[q] for q in v if q in ('ms_per_step','ms_per_iter','solve_ms','cg_iters','cg_iters_per_step','halo_exchanges_per_iter','allreduces_per_iter')})
PY
