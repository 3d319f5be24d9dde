"""One C2 closest-hit pass with chosen knobs (for ncu): env MB, TB, RF, TRIS, S, RAYS, ANY."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_tris = int(os.environ.get("TRIS", 1000000)); s = float(os.environ.get("S", 0.005)); n_rays = int(os.environ.get("RAYS", 1 << 24))
desc = scenes.soup_scene(n_tris, s)
be = backend.B200Backend(); desc.apply(be)
for k, e in (("min_blocks", "MB"), ("tri_batch", "TB"), ("refill_below", "RF")):
    if e in os.environ: be.set_option(k, int(os.environ[e]))
rays = scenes.random_rays(n_rays)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
d_occ = torch.empty(n_rays, dtype=torch.int32, device="cuda")
for _ in range(int(os.environ.get("REPS", 3))):
    if os.environ.get("ANY"): be.trace_any_device(d_rays.data_ptr(), n_rays, d_occ.data_ptr())
    else: be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr())
    print("kernel_ms", be.trace_stats()["kernel_ms"], "Mrays/s", n_rays / be.trace_stats()["kernel_ms"] / 1e3)
