SPP=16 REPS=3 timeout 300 python scripts/profile_render.py > gpurun_out/s4_render23.log 2>&1; grep -o "Msamples/s [0-9.]*" gpurun_out/s4_render23.log; grep "stage ms" gpurun_out/s4_render23.log
timeout 900 python -m pytest tests -m gpu -x -q -k "wavefront or c3 or tile or textur or skinn or render" > gpurun_out/s4_tests23.log 2>&1; tail -3 gpurun_out/s4_tests23.log
