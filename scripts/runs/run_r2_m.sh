for o in wf_split=0 wf_split=1; do
  RFWB200_WF_TRACE=1 WORLD_=8 RANK_=3 SPP=16 REPS=2 STAGES=0 OPTS=$o timeout 300 python scripts/profile_render.py > gpurun_out/r2m_trace_$o.log 2>&1
done
tail -70 gpurun_out/r2m_trace_wf_split=1.log | cut -c1-120
