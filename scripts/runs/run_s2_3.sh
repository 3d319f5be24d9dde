set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests3.log 2>&1; tail -5 gpurun_out/s2_tests3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/s2_bench3.json 2> gpurun_out/s2_bench3.err; cat gpurun_out/s2_bench3.json; tail -3 gpurun_out/s2_bench3.err
