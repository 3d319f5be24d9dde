rm -f gpurun_out/r2n_ab.log
for o in grid_rays_per_thread=0,wf_split=0 grid_rays_per_thread=8,wf_split=0 grid_rays_per_thread=8,wf_split=1 grid_rays_per_thread=16,wf_split=1 grid_rays_per_thread=32,wf_split=1 grid_rays_per_thread=32,wf_split=0 grid_rays_per_thread=64,wf_split=1; do
  AB_WORLD=8 AB_RANK=3 AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2n_ab.log 2>&1
done
for o in grid_rays_per_thread=0,wf_split=0 grid_rays_per_thread=8,wf_split=1 grid_rays_per_thread=32,wf_split=1 grid_rays_per_thread=32,wf_split=0; do AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2n_ab.log 2>&1; done
cut -c82-300 gpurun_out/r2n_ab.log
