for lib in rfw_rs_b200/librfwb200.so build_variants/lib_seg16k.so build_variants/lib_seg64k.so; do RFWB200_LIB=$lib timeout 300 python scripts/exp_c1_flat.py 2>&1 | tail -1; done
for lib in build_variants/lib_seg16k.so build_variants/lib_seg64k.so; do RFWB200_LIB=$lib AB_SKIP_C3=1 timeout 300 python scripts/ab_measure.py 2>&1 | tail -1; done
