mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "elasticity_3d or manufactured or heat_3d or heat_2d or smoke" 2>&1 | tail -3 ) ; 
bash scripts/gpu_ab.sh ab/libpde_base.so ab/libpde_new.so
