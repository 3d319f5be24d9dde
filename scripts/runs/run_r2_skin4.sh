COPIES=9 RFWB200_BUILD_TRACE=1 timeout 600 python scripts/exp_skinning.py 2>&1 | grep -E "largest n = 4672|copies" | tail -4
