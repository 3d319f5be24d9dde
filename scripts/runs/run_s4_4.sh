set -x
mkdir -p gpurun_out
timeout 600 python scripts/tune_c3.py > gpurun_out/s4_tune_c3.log 2>&1; cat gpurun_out/s4_tune_c3.log | tail -20
timeout 600 python -m pytest tests -m gpu -x -q -k "instanced or wavefront or tile or skinn or textur" > gpurun_out/s4_tests4.log 2>&1; tail -3 gpurun_out/s4_tests4.log
