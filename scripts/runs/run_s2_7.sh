set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in librfwb200.so librfwb200_tl7.so librfwb200_tl8.so; do echo "== $lib"; RFWB200_LIB=$PWD/rfw_rs_b200/$lib SPP=16 timeout 200 python scripts/profile_render.py 2>&1 | tail -1 | cut -c1-250; done
TUNE_MB=8 TUNE_TB=6 TUNE_RF=28 timeout 200 python scripts/tune_trace.py 2>&1 | grep -E "^build|min_blocks"
TUNE_TRIS=10000000 TUNE_S=0.002 TUNE_RAYS=4194304 TUNE_MB=8 TUNE_TB=6 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "^build|min_blocks"
python scripts/exp_dynamic.py 2>&1 | tail -1
