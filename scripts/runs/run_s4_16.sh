timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/s4_tests16.log 2>&1; tail -22 gpurun_out/s4_tests16.log
