timeout 600 python scripts/exp_sorted_big.py 2>&1 | tail -9
N_TRIS=5000000 S=0.003 timeout 600 python scripts/exp_sorted_big.py 2>&1 | tail -9
