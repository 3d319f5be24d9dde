timeout 600 python scripts/exp_skinning.py 2>&1 | tail -6
OPTS=build_fused=0 timeout 600 python scripts/exp_skinning.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_skinning.py tests/test_c1_assets.py tests/test_textures.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or splits or instanced or instance_update or small_soups or empty or stress" 2>&1 | tail -3
