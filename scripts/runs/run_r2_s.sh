rm -f gpurun_out/r2s_images.jsonl
RFWB200_IMAGE_LOG=gpurun_out/r2s_images.jsonl timeout 1200 python -u -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --durations=5 > gpurun_out/r2s_pytest.log 2>&1; tail -12 gpurun_out/r2s_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
