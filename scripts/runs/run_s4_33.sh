timeout 600 python -m pytest tests -m gpu -x -q -k "binning" 2>&1 | tail -2
for sr in 0 1; do SORT_RAYS=$sr TUNE_TRIS=10000000 TUNE_S=0.002 TUNE_RAYS=8388608 TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"; done
for sr in 0 1; do SORT_RAYS=$sr TUNE_TRIS=5000000 TUNE_S=0.003 TUNE_RAYS=16777216 TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"; done
SORT_RAYS=1 TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
