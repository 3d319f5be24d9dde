for g in 14 40 70 88; do for f in 1 0; do echo "== grid $g build_fused $f"; GRID=$g OPTS=build_fused=$f timeout 300 python scripts/exp_dynamic.py 2>&1 | tail -1; done; done
