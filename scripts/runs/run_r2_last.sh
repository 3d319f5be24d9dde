timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_skinning.py tests/test_obj.py -x -q -m gpu -k "fused or crowd or skinned or obj or small_soups or instanced_two" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
