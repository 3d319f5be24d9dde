timeout 600 python scripts/exp_skinning.py 2>&1 | tail -6
