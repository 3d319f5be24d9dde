"""Small end-to-end pass for compute-sanitizer: build (BLAS + TLAS + skinned instance + textures), closest / any-hit,
host-streamed tracing, wavefront render with textures, debug view."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes, wire, gltf
desc = scenes.textured_scene(grid=3, subdiv=1, tex_size=16)
be = backend.B200Backend(64, 48); desc.apply(be)
rays = scenes.random_rays(3000, lo=-2.0, hi=2.0)
h = be.trace_closest(rays); o = be.trace_any(rays)
pr = backend.PinnedArray(3000, wire.RAY); ph = backend.PinnedArray(3000, wire.HIT); pr.array[:] = rays
be.trace_closest(pr.array, out=ph.array)
assert np.array_equal(ph.array, h)
view = scenes.camera_view((0, 2.5, -6.0), (0, -0.3, 1.0), 64, 48)
be.render_spp(view, 3, 3); be.render(None, view, 1)
soup = scenes.soup_scene(3000, 0.05); b2 = backend.B200Backend(); soup.apply(b2); b2.trace_closest(scenes.random_rays(5000))
a = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cesium_man.npz"))
sk = gltf.skinned(a, copies=2); b3 = backend.B200Backend(); sk.apply(b3); b3.trace_closest(scenes.random_rays(2000, lo=-1.0, hi=1.0))
print("sanitize smoke ok", int((h["inst"] >= 0).sum()), int(o.sum()))
