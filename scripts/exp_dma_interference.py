"""Does a concurrent PCIe upload (copy engine writing HBM through the L2) slow the traversal kernel?  Plain resident
kernel on C2 alone, then with a 512 MiB pinned H2D / D2H running on another stream at the same time."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n = 1 << 24
desc = scenes.soup_scene(1000000, 0.005)
be = backend.B200Backend(); desc.apply(be)
rays = scenes.random_rays(n)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
hbuf = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); dbuf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
def run(mode, l2p):
    be.set_option("l2_persist", l2p)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            if mode == "h2d": dbuf.copy_(hbuf, non_blocking=True)
            elif mode == "d2h": hbuf.copy_(dbuf, non_blocking=True)
        be.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
        torch.cuda.synchronize()
    print(f"concurrent {mode:5s} l2_persist={l2p}: kernel {best:.2f} ms = {n / best / 1e3:.0f} Mrays/s", flush=True)
for mode in ("none", "h2d", "d2h"):
    for l2p in (0, 1):
        run(mode, l2p)
