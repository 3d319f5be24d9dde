set -x
for lib in librfwb200_base.so librfwb200_n80.so librfwb200.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_MB=8 TUNE_TB=1,3,6 TUNE_RF=28 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "^build|min_blocks|any-hit|simple"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests1.log 2>&1; tail -5 gpurun_out/s3_tests1.log
