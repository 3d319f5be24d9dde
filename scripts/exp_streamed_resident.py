import sys, os; sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n = 1 << 24
desc = scenes.soup_scene(1000000, 0.005)
be = backend.B200Backend(); desc.apply(be)
rays = scenes.random_rays(n)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
for l2p, v in ((1, 0), (0, 0), (1, 2), (0, 2)):
    be.set_option("l2_persist", l2p)
    be.set_option("trace_variant", v)
    best = 1e9
    for _ in range(3):
        be.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
    print("l2_persist", l2p, "variant", v, "kernel_ms", best, "Mrays/s", n / best / 1e3, flush=True)
