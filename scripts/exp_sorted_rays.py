"""Experiment: how much does ray coherence buy the persistent kernel on C2?  Rays are binned on the HOST here (origin
Morton cell + quantised direction) and traced with the unmodified kernel — decides whether a device ray-binning
pre-pass is worth building."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_rays = int(os.environ.get("N_RAYS", 1 << 24))
desc = scenes.soup_scene(int(os.environ.get("N_TRIS", 1000000)), float(os.environ.get("S", 0.005)))
be = backend.B200Backend(); desc.apply(be)
rays = scenes.random_rays(n_rays)
def part(v, bits):
    v = v.astype(np.uint64); out = np.zeros_like(v)
    for b in range(bits): out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out
def keys(rays, ob, db):
    o = np.clip(rays["origin"], 0.0, 0.999999); d = rays["direction"]
    k = np.zeros(len(rays), dtype=np.uint64)
    if ob:
        q = (o * (1 << ob)).astype(np.uint32)
        k = (part(q[:, 0], ob) << np.uint64(2)) | (part(q[:, 1], ob) << np.uint64(1)) | part(q[:, 2], ob)
    if db:
        qd = np.clip(((d + 1.0) * 0.5 * (1 << db)).astype(np.uint32), 0, (1 << db) - 1)
        kd = (part(qd[:, 0], db) << np.uint64(2)) | (part(qd[:, 1], db) << np.uint64(1)) | part(qd[:, 2], db)
        k = (k << np.uint64(3 * db)) | kd
    return k
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
d_occ = torch.empty(n_rays, dtype=torch.int32, device="cuda")
def bench(r, label):
    d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1).copy()).cuda()
    best = 1e9; besta = 1e9
    for _ in range(3):
        be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
    for _ in range(2):
        be.trace_any_device(d_rays.data_ptr(), n_rays, d_occ.data_ptr()); besta = min(besta, be.trace_stats()["kernel_ms"])
    print(f"{label:28s} closest {n_rays / best / 1e3:8.1f} Mrays/s ({best:6.2f} ms)   any {n_rays / besta / 1e3:8.1f} Mrays/s", flush=True)
bench(rays, "unsorted")
for ob, db in [(7, 0), (6, 1), (5, 1), (5, 2), (4, 2), (4, 3), (6, 2), (3, 3), (0, 4), (7, 1)]:
    k = keys(rays, ob, db)
    order = np.argsort(k, kind="stable")
    bench(rays[order], f"origin {ob}b/axis dir {db}b/axis")
for rb in [int(x) for x in os.environ.get("RF", "20,24,28,31").split(",")]:
    be.set_option("refill_below", rb)
    k = keys(rays, 5, 2); bench(rays[np.argsort(k, kind="stable")], f"o5 d2 refill_below {rb}")
