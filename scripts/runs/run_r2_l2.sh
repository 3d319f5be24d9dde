for e in 0 1; do
  if [ $e = 1 ]; then export RFWB200_L2_TRIS=1; O=l2_persist=1; else unset RFWB200_L2_TRIS; O=""; fi
  AB_OPTS=$O timeout 300 python scripts/ab_measure.py 2>&1 | tail -1
  SCENE=c5:10000000 W=3840 H=2160 SPP=16 REPS=3 STAGES=0 OPTS=$O timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | cut -c1-60,300-420
  AB_TRIS=5000000 AB_S=0.003 AB_SKIP_C3=1 AB_OPTS=$O timeout 300 python scripts/ab_measure.py 2>&1 | tail -1
done
