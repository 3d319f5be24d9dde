set -x
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; head -c 300 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
RFWB200_BENCH_STREAMED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --c5-tris 1000000 --c5-spp 4 --c5-frames 1 > gpurun_out/r2g_ncu_list.log 2>&1; tail -1 gpurun_out/r2g_ncu_list.log | cut -c1-200
