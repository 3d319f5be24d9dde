timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_skinning.py tests/test_c1_assets.py tests/test_obj.py tests/test_textures.py -x -q -m gpu -k "fused or small_soups or instanced or skin or c1 or obj or texture or splits or instance_update or empty" 2>&1 | tail -3
COPIES=2,3,4,7,17,65 timeout 600 python scripts/exp_skinning.py 2>&1 | tail -6
timeout 300 python scripts/exp_build_many.py 2>&1 | tail -3
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/r2x3_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok" gpurun_out/r2x3_$tool.log | tail -3
done
