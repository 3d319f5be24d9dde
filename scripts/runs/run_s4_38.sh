timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests38.log 2>&1; tail -3 gpurun_out/s4_tests38.log
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['clocks']['samples'], d['gpu_launches'])"
