set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or small_soups or instanced_two_level or instance_update or empty_and_ragged" 2>&1 | tail -15
for f in 0 1; do echo "== build_fused $f"; BUILD_FUSED=$f timeout 300 python scripts/exp_build_many.py 2>&1 | tail -8; done
