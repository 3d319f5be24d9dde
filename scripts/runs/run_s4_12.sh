set -x
lscpu | grep -E "Socket|NUMA|^CPU\(s\)"; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s4_bench_n8.json 2> gpurun_out/s4_bench_n8.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_n8.json')); print('N=8 value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'pt', d['extra'].get('path_tracing',{}).get('Msamples_per_s'), d['extra'].get('path_tracing',{}).get('gather_ms'))"
tail -3 gpurun_out/s4_bench_n8.err
