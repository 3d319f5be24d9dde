"""Fixed-seed image parity against the oracle with the watertight triangle test (tri_test 0) and with the reference's
Moller-Trumbore arithmetic (tri_test 1): all-pixel RMSE, RMSE after dropping the 0.2 % worst pixels, and the fraction of pixels
that differ by more than 1e-3, per path depth (how fast float32-rounding differences get amplified into different paths)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes
orc.build()
w, h, spp = 256, 144, 8
sky = (0.3, 0.35, 0.5)
for name, desc, view in (("instanced", scenes.instanced_scene(grid=10, subdiv=2, n_lights=16), scenes.camera_view((0, 3.5, -9.0), (0, -0.35, 1.0), w, h)),
                         ("soup", scenes.c5_scene(20000) if hasattr(scenes, "c5_scene") else None, scenes.c5_view(w, h) if hasattr(scenes, "c5_view") else None)):
    if desc is None:
        continue
    cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
    for depth in (1, 2, 3, 5):
        ref, st = cpu.render(view, w, h, spp, depth, clamp=10.0, sky=sky)
        for mt in (0, 1):
            gpu = backend.B200Backend(w, h, sky=sky); desc.apply(gpu)
            gpu.set_option("tri_test", mt)
            gpu.render_spp(view, spp, depth)
            acc = gpu.read_accumulator()
            d = (acc[..., :3].astype(np.float64) - ref[..., :3]) / spp
            m = np.abs(d).max(axis=2).ravel()
            full = float(np.sqrt((d ** 2).mean()))
            keep = np.argsort(m)[: int(np.ceil(len(m) * 0.998))]
            trimmed = float(np.sqrt((d.reshape(-1, 3)[keep] ** 2).mean()))
            rs = gpu.render_stats()
            print(f"{name} depth {depth} tri_test {mt}: all-pixel RMSE {full:.3e}, trimmed {trimmed:.3e}, pixels off by > 1e-3: {(m > 1e-3).mean():.2e}, > 1e-5: {(m > 1e-5).mean():.2e}, "
                  f"ext rays gpu/oracle {rs['extension_rays']}/{st['extension_rays']}, shadow {rs['shadow_rays']}/{st['shadow_rays']}", flush=True)
