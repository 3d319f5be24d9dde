set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests32.log 2>&1; tail -3 gpurun_out/s4_tests32.log
timeout 400 python bench.py > gpurun_out/s4_bench32.json 2> gpurun_out/s4_bench32.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value']); print(d['extra']['path_tracing']); print(d['extra']['bvh_build'], d['extra']['any_hit_Mrays_per_s_per_gpu'], d['extra']['issue_roofline'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_ref32.json 2>> gpurun_out/s4_bench32.err; cut -c1-200 gpurun_out/s4_ref32.json
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 3 -c 1 -o gpurun_out/s4_trace_closest_final -f python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/s4_ncu_full2.log 2>&1; tail -1 gpurun_out/s4_ncu_full2.log
RFWB200_BENCH_STREAMED=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s4_launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s4_ncu_list2.log 2>&1; tail -1 gpurun_out/s4_ncu_list2.log | cut -c1-200
