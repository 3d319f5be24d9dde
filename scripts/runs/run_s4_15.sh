for lib in librfwb200.so librfwb200_n80.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib timeout 400 python scripts/run_configs.py --configs c4 2>&1 | tail -1 | cut -c1-420
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_TRIS=10000000 TUNE_S=0.002 TUNE_RAYS=8388608 TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
done
