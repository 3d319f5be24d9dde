set -x
rm -f gpurun_out/r2f2_images.jsonl
RFWB200_IMAGE_LOG=gpurun_out/r2f2_images.jsonl timeout 1200 python -u -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --durations=3 > gpurun_out/r2f2_pytest.log 2>&1; tail -7 gpurun_out/r2f2_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err; head -c 300 gpurun_out/r2f2_bench.json; tail -3 gpurun_out/r2f2_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f2_ref.json 2>> gpurun_out/r2f2_bench.err; cut -c1-200 gpurun_out/r2f2_ref.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_persistent -s 3 -c 1 -o gpurun_out/r2f2_trace_closest -f python bench.py --steps 2 --warmup 3 --no-extras --no-path-tracing > gpurun_out/r2f2_ncu_full.log 2>&1; tail -1 gpurun_out/r2f2_ncu_full.log | cut -c1-200
RFWB200_BENCH_STREAMED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f2_launches.csv python bench.py --steps 2 --warmup 1 --c5-tris 1000000 --c5-spp 4 --c5-frames 1 > gpurun_out/r2f2_ncu_list.log 2>&1; tail -1 gpurun_out/r2f2_ncu_list.log | cut -c1-200
