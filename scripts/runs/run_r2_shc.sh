timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_skinning.py -x -q -m gpu -k "fused or small_soups or instanced_two or crowd or skinned" 2>&1 | tail -3
COPIES=9 RFWB200_BUILD_TRACE=1 timeout 600 python scripts/exp_skinning.py 2>&1 | grep -E "largest n = 4672|copies" | tail -2
COPIES=2,17,65 timeout 600 python scripts/exp_skinning.py 2>&1 | tail -3
timeout 300 python scripts/exp_build_many.py 2>&1 | tail -3
GRID=14 timeout 300 python scripts/exp_dynamic.py 2>&1 | tail -1
