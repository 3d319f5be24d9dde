timeout 300 python -m pytest tests/test_obj.py -x -q -m gpu 2>&1 | tail -5
SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=1 timeout 600 python scripts/profile_render.py 2>&1 | tail -4
