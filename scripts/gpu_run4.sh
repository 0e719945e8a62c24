mkdir -p gpurun_out
rm -f gpurun_out/r02_modes_v3.jsonl gpurun_out/r02_t4.log
( timeout 600 python -m pytest tests -m gpu -x -q -k "sweep_modes or operator_apply" 2>&1 | tail -2 ) >> gpurun_out/r02_t4.log 2>&1
( PDE_B200_LIB=ab/libpde_tx64ns4.so timeout 600 python -m pytest tests -m gpu -x -q -k "sweep_modes and elasticity" 2>&1 | tail -2 ) >> gpurun_out/r02_t4.log 2>&1
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/r02_modes_v3.jsonl; timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,1,3 >> gpurun_out/r02_modes_v3.jsonl 2>&1; }
run default
for v in tx64ns4 tx64ns2 st6 tx128ns2; do PDE_B200_LIB=ab/libpde_$v.so run $v; done
PDE_B200_E_ZC=32 run zc32
PDE_B200_E_ZC=128 run zc128
PDE_B200_LIB=ab/libpde_tx64ns4.so PDE_B200_E_ZC=32 run tx64ns4_zc32
python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0,3 > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_modes_v3.csv python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0,3 > gpurun_out/ncu4a.log 2>&1
cat gpurun_out/r02_t4.log; python - <<'PY'
import json
for l in open('gpurun_out/r02_modes_v3.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'tag' in d: print('--', d['tag']); continue
    print('  ', d['mode'], d['ms'], d['GBps'])
PY
