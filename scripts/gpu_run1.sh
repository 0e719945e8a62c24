mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "sweep_modes or operator_apply or elasticity_3d or manufactured_solution_large or annihilates" 2>&1 | tail -15 ) > gpurun_out/r02_t1.log 2>&1
for sz in "1280 256 256" "320 64 64"; do
  timeout 300 python scripts/mode_bench.py elasticity $sz >> gpurun_out/r02_modes_new.jsonl 2>&1
  PDE_B200_NO_ELAST3D=1 timeout 300 python scripts/mode_bench.py elasticity $sz >> gpurun_out/r02_modes_old.jsonl 2>&1
done
timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 > gpurun_out/r02_elast_new.jsonl 2>&1
PDE_B200_NO_ELAST3D=1 timeout 300 python scripts/elast_bench.py 1280 256 256 --reps 2 > gpurun_out/r02_elast_old.jsonl 2>&1
cat gpurun_out/r02_t1.log gpurun_out/r02_modes_new.jsonl gpurun_out/r02_modes_old.jsonl gpurun_out/r02_elast_new.jsonl gpurun_out/r02_elast_old.jsonl
