set -x
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" 
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s4_bench_n2.json 2> gpurun_out/s4_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_n2.json')); print('N=2 value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], 'pt', d['extra'].get('path_tracing',{}).get('Msamples_per_s'), d['extra'].get('path_tracing',{}).get('gather_ms'))"
tail -3 gpurun_out/s4_bench_n2.err
