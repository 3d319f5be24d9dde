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
# many small meshes: BLAS builds on the side builder streams; punctual lights + every BSDF lobe; non-finite rays
from rfw_rs_b200 import scenes as _sc
many = _sc.SceneDesc(); many.materials = _sc.material()
for m in range(12):
    many.meshes[m] = _sc.soup(40 + 7 * m, 0.2, seed=100 + m); many.instances[m] = _sc.to_column_major([_sc.trs((m % 4 - 1.5, 0, m // 4 - 1.0))])
for m, n in enumerate((1, 2, 9, 2048, 2049, 5000, 8192)):  # the corner sizes of the fused small build (k_build_small: one CTA per mesh, one launch for all of them)
    many.meshes[12 + m] = _sc.soup(n, 0.2, seed=300 + m); many.instances[12 + m] = _sc.to_column_major([_sc.trs((m - 1.5, 1.0, 0.0))])
b4 = backend.B200Backend(); b4.set_option("build_fused_medium_min", 1); many.apply(b4); r4 = _sc.random_rays(2000, lo=-2.0, hi=2.0); r4["origin"][::7, 0] = np.nan; h4 = b4.trace_closest(r4); b4.trace_any(r4)
b4g = backend.B200Backend(); b4g.set_option("build_fused", 0); many.apply(b4g)  # the general builder on the side streams, same trees
assert b4g.trace_closest(r4).tobytes() == h4.tobytes() and b4g.build_stats()["checksum"] == b4.build_stats()["checksum"]
lob = _sc.lights_and_lobes_scene(grid=3, subdiv=1); b5 = backend.B200Backend(48, 32); lob.apply(b5)
b5.render_spp(_sc.camera_view((0, 2.5, -5.0), (0, -0.4, 1.0), 48, 32, aperture=0.05), 2, 4)
# round 2: the rest of TIntersector, packed hit records (pageable, pinned / streamed, device), the reference's triangle arithmetic
# (TRI_MT builds of the persistent kernels, single- and two-level), the blue-noise sampler, split sub-waves, the tiny-stack build
t = be.intersect_t(rays); t2, dep = be.depth_test(rays)
assert np.array_equal(t >= 0, h["inst"] >= 0) and dep.max() > 0
pk = wire.rays_to_packets4(rays[:2000]); inst4, prim4 = be.intersect4(pk); occ4 = be.occludes4(wire.rays_to_packets4(rays[:2000]))
assert np.array_equal(prim4.ravel(), h["prim"][:2000])
pp = backend.PinnedArray(3000, wire.HIT_PACKED); be.trace_closest_packed(pr.array, out=pp.array); hp = be.trace_closest_packed(rays)
assert np.array_equal(pp.array["prim"], h["prim"]) and np.array_equal(hp["t"], h["t"])
for b in (be, b2):
    b.set_option("tri_test", 1); b.trace_closest(rays); b.trace_any(rays); b.set_option("tri_test", 0)
be.set_option("tri_test", 1); be.reset_accumulator(); be.render_spp(view, 2, 3); be.set_option("tri_test", 0)
be.set_blue_noise(np.random.default_rng(1).integers(0, 256, 65536 * 5).astype(np.uint32)); be.reset_accumulator(); be.render_spp(view, 4, 3); be.set_blue_noise(None)
be.set_option("wf_split", 1); be.reset_accumulator(); be.render_spp(view, 4, 4); be.set_option("wf_split", 0)
b2.set_option("trace_variant", 3); b2.set_option("streamed", 0)
try:
    b2.trace_closest(scenes.random_rays(5000))
except backend.RfwError as e:
    assert "stack overflow" in str(e)
b2.set_option("trace_variant", 0)
print("sanitize smoke ok", int((h["inst"] >= 0).sum()), int(o.sum()))
