"""C3: traversal statistics of the primary rays and of incoherent rays (nodes / triangles / instance entries per ray)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
from oracle import oracle as orc
w, h = 1920, 1080
desc = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
be = backend.B200Backend(); desc.apply(be)
print(be.build_stats())
view = scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, h)
prim = orc.OracleBackend().primary_rays(view, w, h).reshape(-1)
inco = scenes.random_rays(1 << 21, lo=-50.0, hi=50.0); inco["origin"][:, 1] = np.abs(inco["origin"][:, 1]) * 0.05 + 0.05
for name, rays in (("primary", prim), ("incoherent near the ground", inco)):
    d = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda(); dh = torch.empty(len(rays) * 20, dtype=torch.uint8, device="cuda")
    st = be.trace_closest_counted(d.data_ptr(), len(rays), dh.data_ptr())
    be.trace_closest_device(d.data_ptr(), len(rays), dh.data_ptr()); be.trace_closest_device(d.data_ptr(), len(rays), dh.data_ptr())
    ms = be.trace_stats()["kernel_ms"]
    hits = np.frombuffer(dh.cpu().numpy().tobytes(), dtype=wire.HIT)
    print(f"{name}: {len(rays)} rays, nodes/ray {st['nodes_visited'] / len(rays):.1f}, tris/ray {st['tris_tested'] / len(rays):.2f}, instances entered/ray {st['instances_entered'] / len(rays):.2f}, hit rate {(hits['inst'] >= 0).mean():.3f}, {len(rays) / ms / 1e3:.0f} Mrays/s")
