set -x
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests6.log 2>&1; tail -4 gpurun_out/s2_tests6.log
TUNE_MB=8,7 TUNE_TB=8,12,16,20,24 TUNE_BL=2,4,8 TUNE_RF=28 timeout 600 python scripts/tune_trace.py > gpurun_out/s2_tune_spec.log 2>&1; tail -36 gpurun_out/s2_tune_spec.log
