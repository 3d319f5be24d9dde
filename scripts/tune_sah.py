"""Effect of the binned-SAH refinement (treelet size) on build time, SAH cost, traversal steps and throughput."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire, gltf
def dev(a): return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()
def study(label, desc, rays):
    d_rays = dev(rays); n = len(rays)
    d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
    ref = None
    for K in [int(x) for x in os.environ.get("TREELETS", "0,4,8,16,32").split(",")]:
        be = backend.B200Backend(); be.set_option("sah_treelet", K)
        desc.apply(be)
        t0 = time.time(); be.set_option("sah_treelet", K); be.synchronize(); warm_ms = (time.time() - t0) * 1e3   # warm rebuild
        bs = be.build_stats()
        best = 1e9
        for _ in range(3):
            be.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
        h = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT).copy()
        st = be.trace_closest_counted(d_rays.data_ptr(), min(n, 1 << 20), d_hits.data_ptr())
        if ref is None: ref = h
        diff = int(((ref["prim"] != h["prim"]) | (ref["inst"] != h["inst"])).sum())
        print(f"{label} treelet {K:3d}: build {bs['blas_build_ms']:7.2f} ms (sync wall {warm_ms:7.2f}) nodes {bs['blas_nodes']:8d} sah {bs['sah_cost']:7.2f} | nodes/ray {st['nodes_visited']/st['rays']:6.2f} tris/ray {st['tris_tested']/st['rays']:6.2f} | {n/best/1e3:8.1f} Mrays/s | id diffs vs first {diff}", flush=True)
study("C2 soup 1M", scenes.soup_scene(1000000, 0.005), scenes.random_rays(1 << 23))
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
flat = gltf.flatten(asset)
lo, hi = scenes.bounds_of(flat.meshes[0])
r = scenes.random_rays(1 << 22)
r["origin"] = lo + (hi - lo) * r["origin"]   # incoherent rays inside the asset's bounds
study("pica 76k (incoherent)", flat, r)
inst = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
r2 = scenes.random_rays(1 << 22, lo=-50.0, hi=50.0); r2["origin"][:, 1] = np.abs(r2["origin"][:, 1]) * 0.05 + 0.05
study("C3 10k instances", inst, r2)
