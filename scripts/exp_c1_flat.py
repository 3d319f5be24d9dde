"""C1a / C1b primary rays on pica + mixed-scale mesh random rays for the library selected by RFWB200_LIB (A/B of builder knobs on authored geometry)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, gltf, scenes
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
w, h = 1280, 720
flat = gltf.flatten(asset); view = gltf.c1_camera(flat, w, h)
out = [os.path.basename(os.environ.get("RFWB200_LIB", "default"))]
for label, desc in (("C1a flat", flat), ("C1b per-mesh", gltf.per_mesh(asset))):
    be = backend.B200Backend(w, h); desc.apply(be)
    best = 1e9
    for _ in range(8):
        be.cast_primary(view); best = min(best, be.trace_stats()["kernel_ms"])
    bm = []
    for _ in range(4):
        be.set_option("sah_treelet", 8); be.synchronize(); bm.append(be.build_stats()["blas_build_ms"])
    out.append(f"{label} {w*h/best/1e3:7.0f} Mrays/s (warm build {min(bm):.2f} ms)")
desc = scenes.mixed_scale_scene(); be = backend.B200Backend(); desc.apply(be)
rays = scenes.random_rays(1 << 22); d = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda(); hb = torch.empty(len(rays) * 20, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(5):
    be.trace_closest_device(d.data_ptr(), len(rays), hb.data_ptr()); best = min(best, be.trace_stats()["kernel_ms"])
out.append(f"mixed-scale {len(rays)/best/1e3:7.0f} Mrays/s")
print(" | ".join(out), flush=True)
