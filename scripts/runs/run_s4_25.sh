for lib in librfwb200_f0.so librfwb200.so librfwb200_f128.so librfwb200_s256.so; do echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib TUNE_MB=8 TUNE_TB=4 TUNE_RF=28,24 timeout 300 python scripts/tune_trace.py 2>&1 | grep -E "min_blocks|any-hit"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib SPP=16 REPS=3 timeout 300 python scripts/profile_render.py > gpurun_out/s4_render25.log 2>&1; grep -o "Msamples/s [0-9.]*" gpurun_out/s4_render25.log | tail -1; grep "stage ms" gpurun_out/s4_render25.log
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests25.log 2>&1; tail -3 gpurun_out/s4_tests25.log
