set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests4.log 2>&1; tail -5 gpurun_out/s2_tests4.log
python scripts/exp_pcie.py > gpurun_out/s2_pcie.log 2>&1; cat gpurun_out/s2_pcie.log
python scripts/exp_dynamic.py > gpurun_out/s2_dynamic.log 2>&1; tail -3 gpurun_out/s2_dynamic.log
TUNE_MB=8 TUNE_TB=1,2,3,4,6 TUNE_RF=28,30 python scripts/tune_trace.py > gpurun_out/s2_tune_tb.log 2>&1; tail -14 gpurun_out/s2_tune_tb.log
python bench.py --steps 5 --warmup 3 > gpurun_out/s2_bench4.json 2> gpurun_out/s2_bench4.err; cat gpurun_out/s2_bench4.json; tail -3 gpurun_out/s2_bench4.err
# ncu: launch list of the bench command, then one full capture of the closest-hit kernel and of the C3 shade + extend kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1b_bench_launches.csv python bench.py --steps 2 --warmup 3 --pt-spp 4 > gpurun_out/r1b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 3 -c 1 -o gpurun_out/r1b_trace_full python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r1b_ncu_full.log 2>&1
SPP=16 REPS=1 ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 5 -c 1 -o gpurun_out/r1b_shade_full python scripts/profile_render.py > gpurun_out/r1b_ncu_shade.log 2>&1
SPP=16 REPS=1 ncu --set full --clock-control none --import-source on -k regex:ExtendIO -s 6 -c 1 -o gpurun_out/r1b_extend_full python scripts/profile_render.py > gpurun_out/r1b_ncu_extend.log 2>&1
ls -la gpurun_out/*.ncu-rep
