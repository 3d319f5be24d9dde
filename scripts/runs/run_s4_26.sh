timeout 300 python scripts/exp_c3_counts.py 2>&1 | tail -3
