set -x
mkdir -p gpurun_out
for lib in librfwb200_ieee.so librfwb200.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_MB=8 TUNE_TB=4,6,8 TUNE_RF=24,28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit|simple"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib IB=1,6,8,10 timeout 300 python scripts/tune_c3.py 2>&1 | tail -4
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests5.log 2>&1; tail -5 gpurun_out/s4_tests5.log
