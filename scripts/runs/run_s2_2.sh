set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests2.log 2>&1; tail -3 gpurun_out/s2_tests2.log
SPP=16 WAVE_PATHS=1 python scripts/profile_render.py 2>&1 | tail -2
SPP=16 python scripts/profile_render.py 2>&1 | tail -2
SPP=16 WAVE_PATHS=8388608 python scripts/profile_render.py 2>&1 | tail -1
SPP=64 WAVE_PATHS=134217728 python scripts/profile_render.py 2>&1 | tail -1
python scripts/tune_leafcost.py > gpurun_out/s2_leafcost.log 2>&1; cat gpurun_out/s2_leafcost.log
