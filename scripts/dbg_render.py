import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes
desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
w, h, depth = 96, 54, 3
sky = (0.2, 0.2, 0.3)
view = scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h)
cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
def rm(a, b): return float(np.sqrt(np.mean((a[..., :3] - b[..., :3]) ** 2)))
ref0, _ = cpu.render(view, w, h, 1, depth, sky=sky, first_sample=0)
ref1, _ = cpu.render(view, w, h, 1, depth, sky=sky, first_sample=1)
ref01, _ = cpu.render(view, w, h, 2, depth, sky=sky, first_sample=0)
print("oracle self-consistency", rm(ref0 + ref1, ref01))
gpu = backend.B200Backend(w, h, sky=sky); desc.apply(gpu)
gpu.render_spp(view, 1, depth); a0 = gpu.read_accumulator().copy()
gpu.render_spp(view, 1, depth); a01 = gpu.read_accumulator().copy()
print("gpu sample0 vs ref0", rm(a0, ref0), " gpu sample1 (two calls) vs ref1", rm(a01 - a0, ref1))
gpu2 = backend.B200Backend(w, h, sky=sky); desc.apply(gpu2)
gpu2.render_spp(view, 2, depth); b01 = gpu2.read_accumulator().copy()
print("gpu 2spp one call vs ref01", rm(b01, ref01), " vs gpu two calls", rm(b01, a01))
d = np.abs((a01 - a0) - ref1)[..., :3].max(axis=2)
ys, xs = np.nonzero(d > 1e-3)
print("bad pixels", len(ys))
for y, x in list(zip(ys, xs))[:8]:
    print("  px", x, y, "gpu", (a01 - a0)[y, x, :3], "ref", ref1[y, x, :3])
for dd in (1, 2):
    g = backend.B200Backend(w, h, sky=sky); desc.apply(g)
    g.render_spp(view, 1, dd); x0 = g.read_accumulator().copy(); g.render_spp(view, 1, dd); x1 = g.read_accumulator().copy()
    r1, _ = cpu.render(view, w, h, 1, dd, sky=sky, first_sample=1)
    print("depth", dd, "sample1 rmse", rm(x1 - x0, r1))
