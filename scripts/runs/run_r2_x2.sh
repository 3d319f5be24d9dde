for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/r2x2_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok" gpurun_out/r2x2_$tool.log | tail -3
done
