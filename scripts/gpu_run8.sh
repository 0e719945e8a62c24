mkdir -p gpurun_out
CMD2="python scripts/mode_bench.py elasticity 1280 256 256 --reps 2 --modes 0"
$CMD2 > gpurun_out/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_face_rows -s 3 -c 1 -o gpurun_out/r02_face_rows_v5 $CMD2 > gpurun_out/ncu8.log 2>&1
tail -2 gpurun_out/ncu8.log
