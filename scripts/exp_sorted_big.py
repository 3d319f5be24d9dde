"""Ray coherence on scenes whose BVH does not fit the L2 (C4 / C5 sizes): rays binned on the HOST by origin Morton cell (+ direction
octant), traced with the unmodified kernel — decides whether a device ray-sorting pre-pass is worth building for big scenes."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_rays = int(os.environ.get("N_RAYS", 1 << 23))
n_tris, s = int(os.environ.get("N_TRIS", 10000000)), float(os.environ.get("S", 0.002))
desc = scenes.soup_scene(n_tris, s)
be = backend.B200Backend(); desc.apply(be)
print("bvh bytes", be.build_stats()["bvh_bytes"])
rays = scenes.random_rays(n_rays)
def part(v, bits):
    v = v.astype(np.uint64); out = np.zeros_like(v)
    for b in range(bits): out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out
def keys(rays, ob, octant):
    o = np.clip(rays["origin"], 0.0, 0.999999); d = rays["direction"]
    q = (o * (1 << ob)).astype(np.uint32)
    k = (part(q[:, 0], ob) << np.uint64(2)) | (part(q[:, 1], ob) << np.uint64(1)) | part(q[:, 2], ob)
    if octant:
        k = (k << np.uint64(3)) | ((d[:, 0] < 0).astype(np.uint64) << np.uint64(2)) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | (d[:, 2] < 0).astype(np.uint64)
    return k
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
def bench(r, label):
    d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1).copy()).cuda()
    best = 1e9
    for _ in range(3):
        be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
    print(f"{label:32s} closest {n_rays / best / 1e3:8.1f} Mrays/s ({best:6.2f} ms)", flush=True)
bench(rays, "unsorted")
for ob, oc in [(4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (6, 1), (8, 1)]:
    order = np.argsort(keys(rays, ob, oc), kind="stable")
    bench(rays[order], f"origin {ob} bits/axis, octant {oc}")
