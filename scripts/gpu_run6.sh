mkdir -p gpurun_out
rm -f gpurun_out/r02_modes_v5.jsonl
( timeout 900 python -m pytest tests -m gpu -x -q -k "sweep_modes or operator_apply or elasticity_3d" 2>&1 | tail -3 ) > gpurun_out/r02_t6.log 2>&1
timeout 300 python scripts/mode_bench.py elasticity 1280 256 256 >> gpurun_out/r02_modes_v5.jsonl 2>&1
timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 > gpurun_out/r02_elast_v5.jsonl 2>&1
CMD="python scripts/elast_bench.py 1280 256 256 --reps 1"
$CMD > gpurun_out/plain6.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_cfg5_v5.csv $CMD > gpurun_out/ncu6.log 2>&1
cat gpurun_out/r02_t6.log gpurun_out/r02_modes_v5.jsonl gpurun_out/r02_elast_v5.jsonl
