for o in "" "tri_batch_two_level=3" "tri_batch_two_level=6" "inst_batch=3" "inst_batch=10" "refill_below=24" "refill_below=30" "sort_rays=-1" "wave_paths=8388608" "wave_paths=33554432"; do
  echo "== $o"; SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=0 OPTS=$o timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | sed 's/.*Msamples/Msamples/'
done
