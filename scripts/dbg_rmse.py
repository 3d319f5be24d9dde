import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes, wire
sky = (0.3, 0.35, 0.5)
def study(desc, view, w, h, spp, depth, label):
    cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
    gpu = backend.B200Backend(w, h, sky=sky); desc.apply(gpu)
    tot_bad = 0
    for s in range(spp):
        gpu.reset_accumulator(); gpu.set_option("sample_count", s)
        gpu.render_spp(view, 1, depth)
        a = gpu.read_accumulator()[..., :3]
        r, st = cpu.render(view, w, h, 1, depth, sky=sky, first_sample=s)
        r = r[..., :3]
        d = np.abs(a - r).max(axis=2)
        bad = d > 1e-3 * np.maximum(1.0, r.max(axis=2))
        rs = gpu.render_stats()
        tot_bad += bad.sum()
        print(label, "sample", s, "bad pixels", int(bad.sum()), "of", w * h, "ext rays gpu/cpu", rs["extension_rays"], st["extension_rays"], "shadow", rs["shadow_rays"], st["shadow_rays"],
              "rmse", float(np.sqrt(np.mean((a - r) ** 2))))
    print(label, "divergent paths per sample:", tot_bad / (spp * w * h))
w, h = 256, 144
view = scenes.camera_view((0, 3.5, -9.0), (0, -0.35, 1.0), w, h)
study(scenes.instanced_scene(grid=10, subdiv=2, n_lights=16), view, w, h, 4, 5, "subdiv2")
study(scenes.instanced_scene(grid=10, subdiv=2, n_lights=16), view, w, h, 2, 1, "subdiv2-depth1")
study(scenes.instanced_scene(grid=10, subdiv=2, n_lights=16), view, w, h, 2, 2, "subdiv2-depth2")
study(scenes.instanced_scene(grid=10, subdiv=2, n_lights=0), view, w, h, 2, 5, "subdiv2-nolights")
