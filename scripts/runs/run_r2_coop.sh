timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_assets.py -x -q -m gpu -k "fused or small_soups or instanced_two or soup_200k or determinism or splits or c1" 2>&1 | tail -3
timeout 300 python scripts/ab_measure.py 2>&1 | tail -1
COPIES=9 RFWB200_BUILD_TRACE=1 timeout 600 python scripts/exp_skinning.py 2>&1 | grep -E "largest n = 4672|copies" | tail -2
timeout 300 python scripts/exp_c1_flat.py 2>&1 | tail -1
timeout 300 python scripts/exp_build_many.py 2>&1 | tail -3
