timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused" 2>&1 | tail -3
COPIES=2,5,7,8,17,65 timeout 600 python scripts/exp_skinning.py 2>&1 | tail -7
timeout 300 python scripts/exp_build_many.py 2>&1 | tail -3
