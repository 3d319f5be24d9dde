"""One line of numbers for the library build selected by RFWB200_LIB: C2 closest / any-hit Mrays/s (device buffers, best of 5), the
C3 frame (1080p, 16 spp, depth 5, best of 3) and result checksums — the A/B harness for compile-time variants
(scripts/build_variant.sh).  AB_SKIP_C3=1 skips the frame; AB_TRIS / AB_S select another soup."""
import os, sys, zlib; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_tris = int(os.environ.get("AB_TRIS", 1000000)); s = float(os.environ.get("AB_S", 0.005)); n_rays = 1 << 24
name = os.path.basename(os.environ.get("RFWB200_LIB", "default"))
desc = scenes.soup_scene(n_tris, s)
be = backend.B200Backend(); desc.apply(be)
for kv in os.environ.get("AB_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); be.set_option(k, int(v))
build_ms = []
for _ in range(3):
    be.set_option("sah_treelet", 8); be.synchronize(); build_ms.append(be.build_stats()["blas_build_ms"])
rays = scenes.random_rays(n_rays)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
d_occ = torch.empty(n_rays, dtype=torch.int32, device="cuda")
def run(any_hit, reps=5):
    best = 1e9
    for _ in range(reps):
        if any_hit: be.trace_any_device(d_rays.data_ptr(), n_rays, d_occ.data_ptr())
        else: be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr())
        best = min(best, be.trace_stats()["kernel_ms"])
    return n_rays / best / 1e3
run(False, 2)
c, a = run(False), run(True)
h = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT)
crc = zlib.crc32(h["prim"].tobytes()) ^ zlib.crc32(h["t"].tobytes()) ^ zlib.crc32(d_occ.cpu().numpy().tobytes())
line = f"{name:24s} C2({n_tris}) build {min(build_ms):5.2f} ms closest {c:7.1f} any {a:7.1f} Mrays/s crc {crc:08x}"
del be, d_rays, d_hits, d_occ
if not os.environ.get("AB_SKIP_C3"):
    w, hh, spp, depth = 1920, 1080, 16, 5
    d3 = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
    b3 = backend.B200Backend(w, hh, tile_size=64, sky=(0.3, 0.35, 0.5), rank=int(os.environ.get("AB_RANK", 0)), world=int(os.environ.get("AB_WORLD", 1))); d3.apply(b3)
    for kv in os.environ.get("AB_OPTS", "").split(","):
        if "=" in kv:
            k, v = kv.split("="); b3.set_option(k, int(v))
    view = scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, hh)
    best = 1e9
    for _ in range(4):
        b3.reset_accumulator(); b3.render_spp(view, spp, depth); best = min(best, b3.render_stats()["render_ms"])
    acc = b3.read_accumulator()
    line += f" | C3 frame {best:6.2f} ms = {b3.render_stats()['samples'] / best / 1e3:6.1f} Msamples/s crc {zlib.crc32(acc.tobytes()):08x} opts {os.environ.get('AB_OPTS', '')} world {os.environ.get('AB_WORLD', 1)}"
print(line, flush=True)
