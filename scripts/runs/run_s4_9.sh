set -x
for lib in librfwb200.so librfwb200_ld128.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib SPP=16 REPS=2 STAGES=0 timeout 300 python scripts/profile_render.py 2>&1 | tail -1 | cut -c180-420
done
