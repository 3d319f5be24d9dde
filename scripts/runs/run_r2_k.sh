# per-launch durations of one C3 frame (16 spp), whole frame and a 1/8 tile shard, serialised under ncu
for w in 1 8; do
  WORLD_=$w RANK_=3 SPP=16 REPS=1 STAGES=0 OPTS=wf_split=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2k_launches_w$w.csv python scripts/profile_render.py > gpurun_out/r2k_w$w.log 2>&1
  tail -1 gpurun_out/r2k_w$w.log | cut -c1-200
done
