for d in 1 2 3 5; do echo "depth $d"; DEPTH=$d SPP=16 REPS=1 timeout 300 python scripts/profile_render.py 2>&1 | grep -E "stage ms" ; done
