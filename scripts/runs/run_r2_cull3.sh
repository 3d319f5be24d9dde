for i in 1 2; do
RFWB200_LIB=build_variants/lib_precull.so SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=0 timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | sed 's/.*Msamples/precull C5 Msamples/'
SCENE=c5:10000000 W=3840 H=2160 SPP=8 REPS=2 STAGES=0 timeout 300 python scripts/profile_render.py 2>&1 | grep Msamples | tail -1 | sed "s/.*Msamples/new C5 Msamples/"
done
RFWB200_LIB=build_variants/lib_precull.so AB_SKIP_C2=1 timeout 300 python scripts/ab_measure.py 2>&1 | tail -1 | sed 's/.*| C3/precull C3/'
timeout 300 python scripts/ab_measure.py 2>&1 | tail -1 | sed 's/.*| C3/new C3/'
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_skinning.py -x -q -m gpu -k "instanced or instance or skin or wavefront_matches or tintersector or c3_full or c5_full" 2>&1 | tail -3
