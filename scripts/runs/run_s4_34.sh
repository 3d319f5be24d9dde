for lib in librfwb200.so librfwb200_st8.so librfwb200_st16.so; do echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_MB=8 TUNE_TB=4 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib IB=6 TB2=4 RF=28 timeout 300 python scripts/tune_c3.py 2>&1 | tail -1
done
IB=5,7 TB2=3,5 RF=26,30 timeout 300 python scripts/tune_c3.py 2>&1 | tail -8
