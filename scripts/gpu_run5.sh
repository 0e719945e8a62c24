mkdir -p gpurun_out
rm -f gpurun_out/r02_modes_v4.jsonl
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/r02_modes_v4.jsonl; timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,1,3 >> gpurun_out/r02_modes_v4.jsonl 2>&1; }
PDE_B200_LIB=ab/libpde_nomath.so run nomath
PDE_B200_LIB=ab/libpde_nomath.so PDE_B200_E_TOUT=0 run nomath_stg
run default
CMD2="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0"
$CMD2 > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_face_rows -s 3 -c 1 -o gpurun_out/r02_face_rows_v2 $CMD2 > gpurun_out/ncu5.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_modes_v4.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'tag' in d: print('--', d['tag']); continue
    print('  ', d['mode'], d['ms'], d['GBps'])
PY
