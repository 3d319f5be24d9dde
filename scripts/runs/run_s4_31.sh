timeout 900 python scripts/run_configs.py --configs c5 2>&1 | tail -1 | cut -c1-700
