mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/r02_mgpu_check_n$N.log 2>&1
echo "mgpu rc=$?"; grep "\[mgpu\]\|MGPU" gpurun_out/r02_mgpu_check_n$N.log | tail -16
PDE_B200_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_check.py > gpurun_out/r02_mgpu_check_n${N}_nccl.log 2>&1
echo "mgpu nccl rc=$?"; grep "MGPU\|halo" gpurun_out/r02_mgpu_check_n${N}_nccl.log | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
echo "bench rc=$?"
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or multi_gpu" 2>&1 | tail -3 ) > gpurun_out/r02_multi_tool_test_2gpu.log 2>&1; cat gpurun_out/r02_multi_tool_test_2gpu.log; fi
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}.json').read().strip().splitlines()[-1])
print('N=$N weak heat: value',round(d['value'],3),'ms/step',round(d['ms_per_step'],2),'it/step',d['cg_iters_per_step'],'ms/iter',round(d['ms_per_iter'],3), 'halo us',d['halo']['us_per_exchange'] if d['halo'] else None)
e=d['elasticity']; print('   weak elast: iters',e['cg_iters'],'solve',round(e['solve_ms'],1),'ms/iter',round(e['ms_per_iter'],3))
if d['strong']:
  for k,v in d['strong'].items(): print('   strong',k, round(v.get('ms_per_step',v.get('solve_ms')),2), 'ms/iter', round(v['ms_per_iter'],3))
PY
