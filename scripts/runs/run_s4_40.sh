timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests40.log 2>&1; tail -3 gpurun_out/s4_tests40.log
python scripts/exp_dynamic.py 2>&1 | tail -1 | cut -c1-200
