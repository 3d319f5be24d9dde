import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
desc = scenes.soup_scene(1000000, 0.005)
be = backend.B200Backend(); desc.apply(be)
be.set_option("min_blocks", 6); be.set_option("tri_batch", 1); be.set_option("refill_below", 24)
n_rays = 1 << 24
rays = scenes.random_rays(n_rays)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
def t(off, n, reps=3):
    best = 1e9
    for _ in range(reps):
        be.trace_closest_device(d_rays.data_ptr() + off * 32, n, d_hits.data_ptr() + off * 20)
        best = min(best, be.trace_stats()["kernel_ms"])
    return best
for lg in (10, 14, 16, 18, 20, 22, 24):
    n = 1 << lg
    ms = t(0, n)
    print(f"n=2^{lg}: {ms:.3f} ms  {n / ms / 1e3:.1f} Mrays/s")
ch = 1 << 20
times = [t(i * ch, ch, 2) for i in range(16)]
print("16 chunks of 2^20:", ["%.2f" % x for x in times], "sum", sum(times))
# per-ray cost distribution via the counted kernel on small groups: find slow groups of 4096 rays
g = 4096
tt = np.array([t(i * g, g, 1) for i in range(256)])
print("4096-ray groups: min %.3f median %.3f max %.3f ms" % (tt.min(), np.median(tt), tt.max()))
st = be.trace_closest_counted(d_rays.data_ptr(), 1 << 20, d_hits.data_ptr()); print(st)
