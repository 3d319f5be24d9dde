set -x
mkdir -p gpurun_out
SPP=16 REPS=1 timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ExtendIO -s 6 -c 2 -o gpurun_out/s4_extend -f python scripts/profile_render.py > gpurun_out/s4_ncu_extend.log 2>&1; tail -4 gpurun_out/s4_ncu_extend.log | cut -c1-300
SPP=16 REPS=2 timeout 300 python scripts/profile_render.py 2>&1 | tail -2 | cut -c1-400
