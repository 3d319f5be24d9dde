set -x
rm -f gpurun_out/r2h_images.jsonl
RFWB200_IMAGE_LOG=gpurun_out/r2h_images.jsonl timeout 1200 python -u -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --durations=3 > gpurun_out/r2h_pytest.log 2>&1; tail -7 gpurun_out/r2h_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; head -c 300 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_ref.json 2>> gpurun_out/r2h_bench.err; cut -c1-200 gpurun_out/r2h_ref.json
RFWB200_BENCH_STREAMED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-dynamic --c5-tris 1000000 --c5-spp 4 --c5-frames 1 > gpurun_out/r2h_ncu_list.log 2>&1; tail -1 gpurun_out/r2h_ncu_list.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
