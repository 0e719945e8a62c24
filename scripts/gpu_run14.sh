mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus" 2>&1 | tail -15 ) > gpurun_out/r02_multi_tool_test.log 2>&1; cat gpurun_out/r02_multi_tool_test.log
