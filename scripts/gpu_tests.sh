mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
