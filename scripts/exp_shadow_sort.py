"""Would light-sorted shadow rays pay in the connect stage?  C5 scene (10 M-triangle soup + 64 lights): primary hits from the C5 camera ->
next-event-estimation rays toward a random light each (scenes.c4_shadow_rays) -> any-hit time in path order vs stably sorted by light id."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_tris = int(os.environ.get("TRIS", 10_000_000)); w, h = int(os.environ.get("W", 2560)), int(os.environ.get("H", 1440))
desc = scenes.c5_scene(n_tris)
be = backend.B200Backend(w, h); desc.apply(be)
view = scenes.c5_view(w, h)
hits = be.cast_primary(view)
from oracle import oracle as om   # (only for the primary-ray generator of the harness)
rays = om.OracleBackend().primary_rays(view, w, h) if hasattr(om.OracleBackend, "primary_rays") else None
sh, idx = scenes.c4_shadow_rays(desc, rays, hits)
L = desc.area_lights
r0 = scenes.u01(scenes.SEED_LIGHTS + 1, idx)
li = np.minimum((r0 * len(L)).astype(np.int64), len(L) - 1)
def timed(r, label):
    d = torch.from_numpy(r.view(np.uint8).reshape(-1).copy()).cuda(); occ = torch.empty(len(r), dtype=torch.int32, device="cuda")
    best = 1e9
    for _ in range(5):
        be.trace_any_device(d.data_ptr(), len(r), occ.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
    print(f"{label:28s} {len(r)} rays  {best:7.3f} ms  {len(r) / best / 1e3:7.1f} Mrays/s  occluded {float((occ != 0).float().mean()):.3f}", flush=True)
    return occ.cpu().numpy()
a = timed(sh, "path order")
order = np.argsort(li, kind="stable")
b = timed(sh[order], "sorted by light")
assert np.array_equal(a[order], b)
# finer: light id, then 16x16-pixel block of the shading pixel
px = idx % w; py = idx // w
key = (li << 32) | ((py // 16) << 16) | (px // 16)
o2 = np.argsort(key, kind="stable")
timed(sh[o2], "light, then 16x16 pixel block")
rng = np.random.default_rng(1); o3 = rng.permutation(len(sh))
timed(sh[o3], "random order")
