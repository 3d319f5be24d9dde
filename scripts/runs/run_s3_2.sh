set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests2.log 2>&1; tail -5 gpurun_out/s3_tests2.log
RFWB200_PIPE_TRACE=1 CHUNKS=2097152 timeout 300 python scripts/exp_e2e.py > gpurun_out/s3_e2e.log 2>&1; grep -v "^$" gpurun_out/s3_e2e.log | cut -c1-600 | tail -30
