# two GPUs: the C-ABI NCCL test + the bench at N = 2
timeout 300 python -u -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/r2h2_mgpu.log 2>&1; tail -3 gpurun_out/r2h2_mgpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h2_bench_n2.json 2> gpurun_out/r2h2_bench_n2.err; tail -2 gpurun_out/r2h2_bench_n2.err; head -c 200 gpurun_out/r2h2_bench_n2.json
