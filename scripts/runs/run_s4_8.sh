set -x
for lib in librfwb200.so librfwb200_noinl.so; do
  echo "== $lib"
  RFWB200_LIB=$PWD/rfw_rs_b200/$lib SPP=16 REPS=3 timeout 300 python scripts/profile_render.py 2>&1 | tail -3 | cut -c1-60,180-420
done
