"""SURVEY §8 f3, the OBJ half of config C1 ("one obj/gltf mesh from assets/models"): rfw_rs_b200/obj.py against what the reference's
loader glue does with a parsed file (crates/rfw-scene/src/loaders/obj.rs:26-255).  The reference snapshot ships .mtl files but no .obj
(assets/models/{cbox,sponza/sponza,sibenik/sibenik}.mtl), so the geometry fixture tests/golden/box_room.{obj,mtl} is hand-written; it
exercises quads and a pentagon (fan triangulation), v / v/vt/vn / v//vn corners, negative indices, several objects and groups in one
mesh, usemtl switches, an unknown material, a degenerate face, and the `l` / `p` records the loader ignores.  CPU tier: parsing rules
and oracle primary casting; GPU tier: hits and a path-traced image against the oracle."""
import os

import numpy as np
import pytest

from rfw_rs_b200 import obj, scenes
from tests import parity

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def room():
    return obj.load(os.path.join(GOLD, "box_room.obj"))


def test_obj_parsing_rules():
    a = room()
    # 5 room quads + light quad + 5 box quads = 22 triangles, pentagon 3, degenerate 1, the face after the unknown usemtl 1
    assert a.positions.shape == (27, 3, 3)
    assert a.material_names == ["Light", "White", "Red", "Glass"]
    # usemtl: White x3 quads, Red, White, Light, Glass x5 quads, Red pentagon + degenerate, unknown -> the first material (obj.rs:232-240)
    assert a.material_ids.tolist() == [1] * 6 + [2] * 2 + [1] * 2 + [0] * 2 + [3] * 10 + [2] * 4 + [0]
    # fan triangulation of "f 1 4 3 2": (1,4,3), (1,3,2)
    assert np.array_equal(a.positions[0], np.array([[-1, 0, -1], [-1, 0, 1], [1, 0, 1]], np.float32))
    assert np.array_equal(a.positions[1], np.array([[-1, 0, -1], [1, 0, 1], [1, 0, -1]], np.float32))
    # negative indices: the light quad refers to the four vertices just before it
    assert np.array_equal(a.positions[10][:, 1], np.float32([1.98, 1.98, 1.98]))
    # normals / uvs: only the box has both, the pentagon has normals only; everything else zeros
    assert a.normals is not None and a.uvs is not None
    assert np.all(a.normals[:12] == 0) and np.all(np.abs(np.linalg.norm(a.normals[12:22], axis=2) - 1) < 1e-5)
    assert np.array_equal(a.uvs[12], np.float32([[0, 0], [1, 0], [1, 1]])) and np.all(a.uvs[22:] == 0)
    assert a.texture_names["Glass"] == {"map_kd": "glass_albedo.png"}
    t = obj.triangles(a)
    assert len(t) == 26  # the zero-area face is gone
    assert np.array_equal(t["id"], np.arange(26))
    # flat normals where the file gave none, the file's where it did
    assert np.allclose(t["n0"][0], t["normal"][0]) and np.allclose(t["n0"][12], a.normals[12, 0], atol=1e-6)
    assert np.allclose(t["u1"][12], 1.0) and np.allclose(t["v2"][12], 1.0)


def test_mtl_rules_follow_the_loader():
    a = room()
    m = a.materials
    # Light: Ke (0.9, 0.85, 0.7) has every component <= 1 -> x10, then max with Kd (obj.rs:95-99); Ns 0 -> roughness 1
    assert np.allclose(m["color"][0, :3], [9.0, 8.5, 7.0])
    assert (int(m["parameters"][0, 0]) >> 24) == 255
    # White: roughness = 1 - log10(96.078431) / 1000 = 0.99802 -> u8 254; Ks in the specular colour
    assert (int(m["parameters"][1, 0]) >> 24) == int((1 - np.log10(96.078431) / 1000) * 255)
    assert np.allclose(m["specular"][1, :3], 0.5)
    # Glass: transmission 1 - d = 0.75 -> byte 2 of parameters.z; eta 1.45 saturates its u8 slot like into_device_material does
    assert ((int(m["parameters"][3, 2]) >> 16) & 255) == int(0.75 * 255)
    # an emission above 1 is taken as it is; a file without materials gets the red fallback (obj.rs:186-193)
    _, m2, _ = obj.parse_mtl("newmtl L\nKd 0.5 0.5 0.5\nKe 4 0.2 0\n")
    assert np.allclose(m2["color"][0, :3], [4.0, 0.5, 0.5])
    b = obj.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    assert b.material_names == ["<fallback>"] and np.allclose(b.materials["color"][0, :3], [1, 0, 0]) and b.material_ids.tolist() == [0]
    # the reference's own material libraries parse (when the snapshot is there): names and emissive detection
    ref = "/root/reference/assets/models/cbox.mtl"
    if os.path.exists(ref):
        names, mats, _ = obj.parse_mtl(open(ref).read())
        assert names[:4] == ["Light", "DarkGreen", "Khaki", "BloodyRed"]
        assert np.allclose(mats["color"][0, :3], 10.0) and np.allclose(mats["color"][1, :3], [0.0, 0.32, 0.0])
        # the two big libraries of the snapshot: every material parses, the texture maps the loader reads are picked up by name
        # (sponza: 24 diffuse + 22 normal maps through map_Ns — obj.rs:113 reads "map_ns" as the normal map; sibenik: 8 diffuse, 5 bump)
        for rel, n_mats, n_kd, n_norm in (("sponza/sponza.mtl", 25, 24, 22), ("sibenik/sibenik.mtl", 15, 8, 5)):
            names, mats, tex = obj.parse_mtl(open(os.path.join("/root/reference/assets/models", rel), errors="replace").read())
            assert len(names) == n_mats == len(mats) and len(set(names)) == n_mats
            assert sum(1 for t in tex.values() if "map_kd" in t) == n_kd
            assert sum(1 for t in tex.values() if ("map_ns" in t or "map_bump" in t or "bump" in t or "norm" in t)) == n_norm
            assert np.isfinite(mats["color"]).all() and not (mats["color"][:, :3].max(axis=1) > 1.0).any()   # no emissive material in either


def test_obj_scene_casts_on_the_oracle(oracle_mod):
    sc = obj.scene(room())
    assert len(sc.area_lights) == 3 and sorted(sc.meshes[0]["light_id"][sc.meshes[0]["light_id"] >= 0].tolist()) == [0, 1, 2]
    o = oracle_mod.OracleBackend(det_eps=0.0)
    sc.apply(o)
    w, h = 160, 120
    view = scenes.camera_view((0.0, 1.0, 0.95), (0.0, 0.0, -1.0), w, h, fov_deg=110.0)
    rays = o.primary_rays(view, w, h)
    hits = o.trace_closest(rays, mode=oracle_mod.MODE_BVH2)
    brute = o.trace_closest(rays, mode=oracle_mod.MODE_BRUTE)
    assert np.array_equal(hits["prim"], brute["prim"]) and np.array_equal(hits["t"], brute["t"])
    assert (hits["inst"] >= 0).mean() > 0.95          # the camera is inside the room
    seen = set(sc.meshes[0]["mat_id"][hits["prim"][hits["prim"] >= 0]].tolist())
    assert seen == {0, 1, 2, 3}


@pytest.mark.gpu
def test_obj_scene_on_the_gpu_matches_the_oracle(oracle_mod):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfw_rs_b200 import backend

    sc = obj.scene(room())
    w, h, spp, depth = 160, 120, 16, 5
    view = scenes.camera_view((0.0, 1.0, 0.95), (0.0, 0.0, -1.0), w, h, fov_deg=110.0)
    gpu = backend.B200Backend(w, h)
    sc.apply(gpu)
    cpu = oracle_mod.OracleBackend(det_eps=0.0)
    sc.apply(cpu)
    rays = cpu.primary_rays(view, w, h)
    ref = cpu.trace_closest(rays, mode=oracle_mod.MODE_BVH2)
    parity.compare_hits(rays, gpu.trace_closest(rays), ref, parity.lookup_from_desc(sc), "box_room.obj")
    assert np.array_equal(gpu.trace_any(rays), cpu.trace_any(rays, mode=oracle_mod.MODE_BVH2))
    # path-traced: glass box (transmission), emissive quad, MIS — radiance RMSE at the bar of tests/test_gpu_parity.py::check_image
    gpu.render_spp(view, spp, depth)
    acc = gpu.read_accumulator() / spp
    img, _ = cpu.render(view, w, h, spp, depth, clamp=10.0, sky=(0.0, 0.0, 0.0))
    img = img / spp
    assert np.isfinite(acc).all() and img[..., :3].mean() > 0.01
    d = np.abs(acc[..., :3].astype(np.float64) - img[..., :3].astype(np.float64)).max(axis=2).ravel()
    keep = np.argsort(d)[: int(np.ceil(len(d) * 0.998))]
    sq = ((acc[..., :3].astype(np.float64) - img[..., :3].astype(np.float64)) ** 2).reshape(-1, 3)
    assert float(np.sqrt(sq[keep].mean())) <= 1e-3          # 99.8 % of the pixels (see check_image for why not all)
    assert float(np.sqrt(sq.mean())) <= 1e-2
