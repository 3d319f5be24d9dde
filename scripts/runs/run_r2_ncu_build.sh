COPIES=65 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_build_small -s 6 -c 1 -o gpurun_out/r2_build_small512 -f python scripts/exp_skinning.py > gpurun_out/r2_ncu_build512.log 2>&1; tail -2 gpurun_out/r2_ncu_build512.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_build_small -s 4 -c 1 -o gpurun_out/r2_build_small256 -f python scripts/exp_build_many.py > gpurun_out/r2_ncu_build256.log 2>&1; tail -2 gpurun_out/r2_ncu_build256.log | cut -c1-200
ls -la gpurun_out/r2_build_small*.ncu-rep
