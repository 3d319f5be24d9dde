set -x
mkdir -p gpurun_out
for lib in librfwb200.so librfwb200_tl8.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib IB=4,6,8 TB2=4,6 timeout 300 python scripts/tune_c3.py 2>&1 | tail -6
done
timeout 900 python -m pytest tests -m gpu -x -q -k "instanced or wavefront or tile or skinn or textur or instance" > gpurun_out/s4_tests6.log 2>&1; tail -3 gpurun_out/s4_tests6.log
