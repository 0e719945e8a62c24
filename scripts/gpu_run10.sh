mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/r02_mgpu_check_n$N.log 2>&1
echo "mgpu rc=$?"; grep "\[mgpu\]\|MGPU\|Error\|error" gpurun_out/r02_mgpu_check_n$N.log | tail -25
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n${N}_a.json 2> gpurun_out/r02_bench_n${N}_a.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n${N}_a.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','cg_iters_per_step','ms_per_iter','true_relres','roofline_step','halo'): print(k, d.get(k))
e=d['elasticity']; print('elast', {k:e[k] for k in e if k not in ('sweeps','roofline','workload')})
print('strong', json.dumps(d['strong'])[:1500])
PY
