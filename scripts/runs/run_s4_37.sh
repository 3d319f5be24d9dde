timeout 300 python scripts/exp_interactive.py 2>&1 | tail -5
