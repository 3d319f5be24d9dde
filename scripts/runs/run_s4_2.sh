set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests2.log 2>&1; tail -5 gpurun_out/s4_tests2.log
timeout 400 python bench.py > gpurun_out/s4_bench2.json 2> gpurun_out/s4_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench2.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['extra'].get('path_tracing',{}).get('Msamples_per_s'), d['extra'].get('bvh_build'))"
RFWB200_BENCH_STREAMED=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s4_ncu_list.log 2>&1; tail -2 gpurun_out/s4_ncu_list.log | cut -c1-300
timeout 600 python scripts/run_configs.py --configs c1,c4 > gpurun_out/s4_configs.log 2>&1; tail -5 gpurun_out/s4_configs.log | cut -c1-900
python __graft_entry__.py smoke 2>&1 | tail -2
