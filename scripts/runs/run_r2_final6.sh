set -x
timeout 1200 python -u -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/r2k2_pytest.log 2>&1; tail -3 gpurun_out/r2k2_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k2_bench.json 2> gpurun_out/r2k2_bench.err; head -c 200 gpurun_out/r2k2_bench.json; tail -3 gpurun_out/r2k2_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
