"""Spatial splits (option split_budget, tri_split.h) on a mesh with mixed triangle scales: node visits / triangle tests per ray, Mrays/s and
build time against the budget; hits must not change."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire, gltf
n = 1 << 22
rays = scenes.random_rays(n)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
def measure(desc, label, budgets=(0, 10, 30, 60, 100)):
    ref = None
    for b in budgets:
        be = backend.B200Backend(); be.set_option("split_budget", b); desc.apply(be)
        be.set_option("sah_treelet", 8); t0 = time.perf_counter(); be.synchronize(); sync_ms = (time.perf_counter() - t0) * 1e3
        best = 1e9
        for _ in range(3):
            be.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
        h = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT).copy()
        st = be.trace_closest_counted(d_rays.data_ptr(), 1 << 20, d_hits.data_ptr())
        bs = be.build_stats()
        if ref is None: ref = h
        same = bool(np.array_equal(h["prim"], ref["prim"]) and np.array_equal(h["t"], ref["t"]))
        print(f"{label} budget {b:3d} %: {n / best / 1e3:7.1f} Mrays/s, nodes/ray {st['nodes_visited'] / st['rays']:6.2f}, tris/ray {st['tris_tested'] / st['rays']:5.2f}, "
              f"wide nodes {bs['blas_nodes']}, bvh {bs['bvh_bytes'] / 1e6:.1f} MB, sah {bs['sah_cost']:.1f}, rebuild {sync_ms:.1f} ms, hits identical {same}", flush=True)
measure(scenes.mixed_scale_scene(), "mixed scales (100k small + 300 needles + 30 big)")
measure(scenes.soup_scene(1000000, 0.005), "C2 soup", budgets=(0, 30))
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
flat = gltf.flatten(asset)
lo = flat.meshes[0]["vertex0"].min(axis=0); hi = flat.meshes[0]["vertex0"].max(axis=0)
rays2 = scenes.random_rays(n, lo=0.0, hi=1.0); rays2["origin"] = lo + rays2["origin"] * (hi - lo)
d_rays.copy_(torch.from_numpy(rays2.view(np.uint8).reshape(-1).copy()))
measure(flat, "pica flattened (76k triangles)", budgets=(0, 30, 100))
