SPP=16 REPS=3 timeout 300 python scripts/profile_render.py > gpurun_out/s4_render20.log 2>&1; grep -o "Msamples/s [0-9.]*" gpurun_out/s4_render20.log; grep "stage ms" gpurun_out/s4_render20.log
