timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or small_soups or instanced or instance_update or empty or stress" 2>&1 | tail -3
timeout 600 python scripts/exp_skinning.py 2>&1 | tail -4
timeout 300 python scripts/exp_build_many.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_skinning.py tests/test_c1_assets.py tests/test_textures.py tests/test_obj.py -x -q -m gpu 2>&1 | tail -3
