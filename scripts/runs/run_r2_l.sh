rm -f gpurun_out/r2l_ab.log
for mc in 8 32; do for o in wf_split=0 wf_split=1; do CUDA_DEVICE_MAX_CONNECTIONS=$mc AB_WORLD=8 AB_RANK=3 AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2l_ab.log 2>&1; done; done
for o in wf_split=0 wf_split=1; do CUDA_DEVICE_MAX_CONNECTIONS=32 AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2l_ab.log 2>&1; done
cut -c60-300 gpurun_out/r2l_ab.log
