mkdir -p gpurun_out
rm -f gpurun_out/r02_modes_v7.jsonl
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/r02_modes_v7.jsonl; timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,3 >> gpurun_out/r02_modes_v7.jsonl 2>&1; }
run default
PDE_B200_LIB=ab/libpde_face3.so run face3
PDE_B200_LIB=ab/libpde_face4.so run face4
python - <<'PY'
import json
for l in open('gpurun_out/r02_modes_v7.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'tag' in d: print('--', d['tag']); continue
    print('  ', d['mode'], d['ms'], d['GBps'])
PY
# ncu: DRAM traffic + full metrics of the final elasticity apply kernel and the heat post2 kernel
CMD2="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0"
$CMD2 > gpurun_out/plain15.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_elast3d -s 3 -c 1 -o gpurun_out/r02_elast_apply_final $CMD2 > gpurun_out/ncu15a.log 2>&1
CMD3="python bench.py --steps 2 --warmup 3 --no-elasticity --no-configs --no-cpu"
$CMD3 > gpurun_out/plain15b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_heat_post2 -s 20 -c 1 -o gpurun_out/r02_heat_post2_final $CMD3 > gpurun_out/ncu15b.log 2>&1
tail -2 gpurun_out/ncu15a.log gpurun_out/ncu15b.log
