mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-elasticity --no-configs --no-cpu"
$CMD > gpurun_out/plainf2a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_heat_cfg4.csv $CMD > gpurun_out/ncuf2a.log 2>&1
CMD="python scripts/elast_bench.py 1280 256 256 --reps 1"
$CMD > gpurun_out/plainf2b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_elast_cfg5.csv $CMD > gpurun_out/ncuf2b.log 2>&1
CMD3="python bench.py --steps 2 --warmup 3 --no-elasticity --no-configs --no-cpu"
$CMD3 > gpurun_out/plainf2c.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:k_heat_post2<\(int\)64, \(int\)4, \(bool\)0' -s 4 -c 1 -o gpurun_out/r02_heat_post2_final $CMD3 > gpurun_out/ncuf2c.log 2>&1
tail -2 gpurun_out/ncuf2c.log
