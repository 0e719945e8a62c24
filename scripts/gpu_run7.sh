mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "sweep_modes or operator_apply" 2>&1 | tail -3 ) > gpurun_out/r02_t7.log 2>&1
timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 --modes 0,3 > gpurun_out/r02_modes_v6.jsonl 2>&1
CMD="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0,3"
$CMD > gpurun_out/plain7.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_modes_v6.csv $CMD > gpurun_out/ncu7.log 2>&1
cat gpurun_out/r02_t7.log gpurun_out/r02_modes_v6.jsonl
