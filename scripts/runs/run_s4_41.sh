timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests41.log 2>&1; tail -2 gpurun_out/s4_tests41.log
