for tool in racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/r2x4_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok" gpurun_out/r2x4_$tool.log | tail -2
done
