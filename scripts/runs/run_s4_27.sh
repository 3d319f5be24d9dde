SPP=16 REPS=3 timeout 300 python scripts/profile_render.py > gpurun_out/s4_render27.log 2>&1; grep -o "Msamples/s [0-9.]*" gpurun_out/s4_render27.log; grep "stage ms" gpurun_out/s4_render27.log
timeout 300 python scripts/exp_c3_counts.py 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests27.log 2>&1; tail -3 gpurun_out/s4_tests27.log
