timeout 300 python scripts/exp_sparse.py 2>&1 | tail -6
TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
SPP=16 REPS=3 STAGES=0 timeout 300 python scripts/profile_render.py 2>&1 | grep -o "Msamples/s [0-9.]*"
echo "== without the tick bound"
RFWB200_LIB=$PWD/rfw_rs_b200/librfwb200_notick.so timeout 300 python scripts/exp_sparse.py 2>&1 | tail -6
