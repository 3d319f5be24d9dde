set -x
mkdir -p gpurun_out
IB=6 TB2=4 timeout 300 python scripts/tune_c3.py 2>&1 | tail -1
SPP=16 REPS=2 timeout 300 python scripts/profile_render.py 2>&1 | tail -3 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests7.log 2>&1; tail -3 gpurun_out/s4_tests7.log
