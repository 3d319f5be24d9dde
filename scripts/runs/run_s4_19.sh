SPP=16 REPS=3 timeout 300 python scripts/profile_render.py 2>&1 | tail -3 | cut -c1-50,250-420
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests19.log 2>&1; tail -3 gpurun_out/s4_tests19.log
