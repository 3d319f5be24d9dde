set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; tail -3 gpurun_out/s2_tests.log
TUNE_MB=8 TUNE_TB=1 TUNE_RF=28 python scripts/tune_trace.py > gpurun_out/s2_tune_stackfix.log 2>&1; tail -4 gpurun_out/s2_tune_stackfix.log
python scripts/exp_sorted_rays.py > gpurun_out/s2_sorted.log 2>&1; cat gpurun_out/s2_sorted.log
