timeout 800 python scripts/exp_shadow_sort.py 2>&1 | tail -8
