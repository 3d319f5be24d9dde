RFWB200_PIPE_TRACE=1 timeout 300 python scripts/exp_e2e_trace.py 2>&1 | tail -14 | cut -c1-700
