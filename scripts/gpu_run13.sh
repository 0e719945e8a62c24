mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 10 --warmup 3 --no-elasticity --no-configs --no-cpu > gpurun_out/tmp13.json 2>gpurun_out/tmp13.err; python -c "
import json; d=json.loads(open('gpurun_out/tmp13.json').read()); print('$1', d['ms_per_step'], d['cg_iters_per_step'], d['roofline_step']['frac'])"; }
run base
PDE_B200_P2_YSB=3 run ysb3
PDE_B200_P2_YSB=2 run ysb2
PDE_B200_P2_ZC=128 run zc128
PDE_B200_P2_ZC=32 run zc32
PDE_B200_P2_CW=32 run cw32
