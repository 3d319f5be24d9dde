timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or small_soups or instanced_two" 2>&1 | tail -3
COPIES=9 RFWB200_BUILD_TRACE=1 timeout 600 python scripts/exp_skinning.py 2>&1 | grep -E "largest n = 4672|copies" | tail -2
COPIES=2,7,17,65 OPTS=build_fused_medium_min=1 timeout 600 python scripts/exp_skinning.py 2>&1 | tail -4
OPTS=build_fused_medium_min=1 timeout 300 python scripts/exp_build_many.py 2>&1 | tail -3
