for lib in rfw_rs_b200/librfwb200.so build_variants/lib_seg512.so build_variants/lib_seg1024.so rfw_rs_b200/librfwb200.so; do RFWB200_LIB=$lib timeout 300 python scripts/exp_c1_flat.py 2>&1 | tail -1; done
for lib in rfw_rs_b200/librfwb200.so build_variants/lib_seg512.so; do RFWB200_LIB=$lib AB_TRIS=5000000 AB_S=0.003 AB_SKIP_C3=1 timeout 300 python scripts/ab_measure.py 2>&1 | tail -1; done
