mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "sweep_modes" 2>&1 | tail -5 ) > gpurun_out/r02_t2.log 2>&1
CMD="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_modes.csv $CMD > gpurun_out/ncu2a.log 2>&1
CMD2="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0,3"
$CMD2 > gpurun_out/plain2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_elast3d -s 3 -c 1 -o gpurun_out/r02_elast_apply_v1 $CMD2 > gpurun_out/ncu2b.log 2>&1
cat gpurun_out/r02_t2.log; tail -3 gpurun_out/ncu2a.log gpurun_out/ncu2b.log
