timeout 600 python -m pytest tests/test_skinning.py -x -q -m gpu 2>&1 | tail -5
