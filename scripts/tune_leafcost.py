"""Sweep the SAH leaf parameters (cost of a triangle test relative to a node visit, max triangles per leaf slot) on
C2: triangle tests run at ~12 of 32 lanes in the persistent kernel, node tests at ~24, so the cost ratio the DP
should use is a property of the kernel, not of the textbook."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_rays = int(os.environ.get("N_RAYS", 1 << 23))
def dev(a): return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()
def study(label, desc, rays):
    d_rays = dev(rays); n = len(rays)
    d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda"); d_occ = torch.empty(n, dtype=torch.int32, device="cuda")
    be = backend.B200Backend(); desc.apply(be)
    ref = None
    for pmax in [int(x) for x in os.environ.get("PMAX", "3,2,1").split(",")]:
        for cp in [int(x) for x in os.environ.get("CPRIM", "150,300,500,800,1200,2000").split(",")]:
            be.set_option("sah_pmax", pmax); be.set_option("sah_c_prim_milli", cp); be.synchronize()
            bs = be.build_stats()
            best = 1e9; besta = 1e9
            for _ in range(3):
                be.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
            h = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT).copy()
            for _ in range(2):
                be.trace_any_device(d_rays.data_ptr(), n, d_occ.data_ptr()); besta = min(besta, be.trace_stats()["kernel_ms"])
            st = be.trace_closest_counted(d_rays.data_ptr(), min(n, 1 << 20), d_hits.data_ptr())
            if ref is None: ref = h
            diff = int(((ref["prim"] != h["prim"]) | (ref["inst"] != h["inst"])).sum())
            print(f"{label} pmax {pmax} c_prim {cp/1000:5.2f}: build {bs['blas_build_ms']:6.2f} ms nodes {bs['blas_nodes']:8d} bvh {bs['bvh_bytes']/1e6:6.1f} MB | nodes/ray {st['nodes_visited']/st['rays']:6.2f} "
                  f"tris/ray {st['tris_tested']/st['rays']:6.2f} | closest {n/best/1e3:8.1f} any {n/besta/1e3:8.1f} Mrays/s | id diffs {diff}", flush=True)
study("C2", scenes.soup_scene(1000000, 0.005), scenes.random_rays(n_rays))
if os.environ.get("WITH_C4", "1") == "1":
    os.environ["PMAX"] = os.environ.get("PMAX4", "3,1"); os.environ["CPRIM"] = os.environ.get("CPRIM4", "300,800,2000")
    study("C4-like 5M", scenes.soup_scene(5000000, 0.003), scenes.random_rays(n_rays))
