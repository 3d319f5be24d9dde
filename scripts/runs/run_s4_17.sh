for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/s4_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "SUMMARY|sanitize smoke ok" gpurun_out/s4_$tool.log | tail -2
done
