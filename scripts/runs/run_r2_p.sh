RFWB200_IMAGE_LOG=gpurun_out/r2p_images.jsonl timeout 900 python -u -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --durations=5 > gpurun_out/r2p_pytest.log 2>&1; tail -12 gpurun_out/r2p_pytest.log
AB_SKIP_C3=1 timeout 200 python scripts/ab_measure.py
