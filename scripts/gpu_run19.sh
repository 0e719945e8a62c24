mkdir -p gpurun_out
CMD="python scripts/elast_bench.py 1280 256 256 --reps 1"
$CMD > gpurun_out/plain19.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_cfg5_v6.csv $CMD > gpurun_out/ncu19.log 2>&1
tail -1 gpurun_out/plain19.log | cut -c1-300
