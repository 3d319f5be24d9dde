python scripts/exp_pcie.py 2>&1 | tail -4
for i in 1 2; do
for lib in build_variants/lib_old.so rfw_rs_b200/librfwb200.so; do
  echo "== $lib"; RFWB200_LIB=$lib CHUNKS=2097152 timeout 300 python scripts/exp_e2e.py 2>&1 | grep "l2_persist=0 streamed=1"
done; done
