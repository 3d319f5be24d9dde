for i in 1 2; do
RFWB200_LIB=build_variants/lib_precull.so SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=0 timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | sed 's/.*Msamples/precull Msamples/'
for o in "instance_box_cull=1" "instance_box_cull=0"; do
  SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=0 OPTS=$o timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | sed "s/.*Msamples/$o Msamples/"
done; done
