mkdir -p gpurun_out
CMD3="python bench.py --steps 1 --warmup 3 --no-elasticity --no-configs --no-cpu"
$CMD3 > gpurun_out/plain_p2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_heat_post2 -s 24 -c 12 -o gpurun_out/r02_heat_post2_final $CMD3 > gpurun_out/ncu_p2.log 2>&1
tail -n 2 gpurun_out/ncu_p2.log
