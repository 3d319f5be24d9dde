set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests1.log 2>&1; tail -5 gpurun_out/s4_tests1.log
timeout 200 python scripts/exp_build_many.py > gpurun_out/s4_build_many.log 2>&1; tail -14 gpurun_out/s4_build_many.log
timeout 400 python bench.py > gpurun_out/s4_bench1.json 2> gpurun_out/s4_bench1.err; cut -c1-1500 gpurun_out/s4_bench1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_ref1.json 2>> gpurun_out/s4_bench1.err; cut -c1-600 gpurun_out/s4_ref1.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 3 -c 1 -o gpurun_out/s4_trace_closest -f python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/s4_ncu_full.log 2>&1; tail -3 gpurun_out/s4_ncu_full.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s4_ncu_list.log 2>&1; tail -2 gpurun_out/s4_ncu_list.log
ls -la gpurun_out
