mkdir -p gpurun_out
N=8
for HB in 148 48 24; do
PDE_B200_HALO_BLOCKS=$HB timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-elasticity --no-configs --no-cpu > gpurun_out/hb_$HB.json 2> gpurun_out/hb_$HB.err
python - <<PY
import json
d=json.loads(open('gpurun_out/hb_$HB.json').read().strip().splitlines()[-1])
print('HALO_BLOCKS=$HB halo us', round(d['halo']['us_per_exchange'],2), 'weak ms/iter', round(d['ms_per_iter'],3), 'strong ms/iter', round(d['strong']['heat']['ms_per_iter'],3))
PY
done
