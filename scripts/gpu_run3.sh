mkdir -p gpurun_out
rm -f gpurun_out/r02_modes_v2.jsonl
for ys in 2 3 4; do for tout in 1 0; do
  export PDE_B200_E_YS=$ys PDE_B200_E_TOUT=$tout
  ( timeout 600 python -m pytest tests -m gpu -x -q -k "sweep_modes and elasticity" 2>&1 | tail -2 ) >> gpurun_out/r02_t3.log 2>&1
  timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 >> gpurun_out/r02_modes_v2.jsonl 2>&1
done; done
unset PDE_B200_E_YS PDE_B200_E_TOUT
python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_modes_v2.csv python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 > gpurun_out/ncu3a.log 2>&1
cat gpurun_out/r02_t3.log; python - <<'PY'
import json
for l in open('gpurun_out/r02_modes_v2.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['env'].get('PDE_B200_E_YS'), d['env'].get('PDE_B200_E_TOUT'), d['mode'], d['ms'], d['GBps'])
PY
