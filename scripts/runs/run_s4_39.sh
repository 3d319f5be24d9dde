timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests39.log 2>&1; tail -3 gpurun_out/s4_tests39.log
SPP=16 REPS=3 timeout 300 python scripts/profile_render.py > gpurun_out/s4_render39.log 2>&1; grep -o "Msamples/s [0-9.]*" gpurun_out/s4_render39.log | tail -2; grep "stage ms" gpurun_out/s4_render39.log
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
