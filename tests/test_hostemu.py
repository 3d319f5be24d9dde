"""CPU-tier check of the PRODUCT's builder / traversal logic (rfw_rs_b200/csrc/{bvh_build.h,traverse.h})
compiled for the host by tests/hostemu, against the oracle.  Not a product path: see tests/hostemu/emu.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from rfw_rs_b200 import scenes, wire
from tests import parity

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostemu"), "-s"])
    L = C.CDLL(os.path.join(HERE, "hostemu", "libemu.so"))
    L.emu_create.restype = C.c_void_p
    L.emu_destroy.argtypes = [C.c_void_p]
    L.emu_set_mesh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.emu_set_instances.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.emu_build.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
    L.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_validate.argtypes = [C.c_void_p, C.c_uint32]
    L.emu_validate.restype = C.c_int
    L.emu_sah.argtypes = [C.c_void_p, C.c_uint32]
    L.emu_sah.restype = C.c_float
    return L


class Emu:
    def __init__(self, L, desc):
        self.L = L
        self.h = C.c_void_p(L.emu_create())
        self.keep = []
        for mid, t in desc.meshes.items():
            t = np.ascontiguousarray(t); self.keep.append(t)
            L.emu_set_mesh(self.h, mid, t.ctypes.data, len(t))
        for mid, m in desc.instances.items():
            m = np.ascontiguousarray(m, np.float32); self.keep.append(m)
            L.emu_set_instances(self.h, mid, m.ctypes.data, len(m))
        self.stats = np.zeros(3, np.uint64)
        L.emu_build(self.h, 1.0, 0.3, 3, self.stats.ctypes.data)

    def trace(self, rays):
        hits = np.empty(len(rays), wire.HIT); occ = np.empty(len(rays), np.uint32); ctr = np.zeros(3, np.uint64)
        self.L.emu_trace(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, occ.ctypes.data, ctr.ctypes.data)
        return hits, occ, ctr

    def __del__(self):
        self.L.emu_destroy(self.h)


@pytest.mark.parametrize("n_tris", [1, 2, 3, 4, 9, 100, 5000])
def test_collapse_structure_and_parity_soup(emu, oracle_mod, n_tris):
    desc = scenes.soup_scene(n_tris, 0.3 if n_tris < 200 else 0.04)
    e = Emu(emu, desc)
    assert emu.emu_validate(e.h, 0) == 0
    o = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(o)
    rays = scenes.random_rays(20000 if n_tris >= 100 else 2000)
    hits, occ, ctr = e.trace(rays)
    ref = o.trace_closest(rays)
    parity.compare_hits(rays, hits, ref, parity.lookup_from_desc(desc), f"soup{n_tris}")
    ref_occ = o.trace_any(rays)
    assert (occ != ref_occ).sum() <= 2
    if n_tris >= 5000:
        assert (ref["inst"] >= 0).mean() > 0.3
        assert ctr[0] / len(rays) < 60  # wide nodes per ray stays sane


def test_two_level_parity(emu, oracle_mod):
    desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
    e = Emu(emu, desc)
    for mid in desc.meshes:
        assert emu.emu_validate(e.h, mid) == 0
    o = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(o)
    rays = scenes.random_rays(30000, lo=-4.0, hi=4.0)
    rays["origin"][:, 1] = np.abs(rays["origin"][:, 1]) * 0.5 + 0.05
    hits, occ, ctr = e.trace(rays)
    ref = o.trace_closest(rays)
    nbad = parity.compare_hits(rays, hits, ref, parity.lookup_from_desc(desc), "instanced")
    assert (ref["inst"] >= 0).mean() > 0.3
    assert (occ != o.trace_any(rays)).sum() <= 2
    assert int(e.stats[2]) == 36 + 1 + 4


def test_axis_aligned_and_degenerate_rays(emu, oracle_mod):
    # axis-parallel rays (zero direction components -> inf reciprocals) against an axis-aligned quad
    desc = scenes.SceneDesc()
    desc.meshes[0] = scenes.quad((0, 0, 0), (0, 1, 0), 2.0, 2.0)
    desc.instances[0] = scenes.to_column_major([scenes.identity()])
    desc.materials = scenes.material()
    e = Emu(emu, desc)
    o = oracle_mod.OracleBackend(); desc.apply(o)
    rays = np.zeros(5, wire.RAY)
    rays["origin"] = [(0.3, 1, 0.2), (0.3, -1, 0.2), (0.3, 1, 0.2), (5, 1, 0), (-0.5, 2, -0.5)]
    rays["direction"] = [(0, -1, 0), (0, 1, 0), (0, 1, 0), (0, -1, 0), (0, -1, 0)]
    rays["tmin"] = 1e-4; rays["tmax"] = 1e26
    hits, occ, _ = e.trace(rays)
    ref = o.trace_closest(rays, mode=oracle_mod.MODE_BRUTE)
    assert np.array_equal(hits["inst"], ref["inst"]) and np.array_equal(hits["prim"], ref["prim"])
    assert list(hits["inst"]) == [0, 0, -1, -1, 0]
    assert np.allclose(hits["t"][[0, 1, 4]], [1, 1, 2])


def test_non_finite_rays_retire_as_misses(emu, oracle_mod):
    """NaN / infinite ray components: every comparison of the reference's tests is false for them (a miss).  The wide-node
    test drops NaN slab operands, so without the guard of traverse.h such a ray would be accepted by every box and walk the
    whole tree (a batch of them would stall the device for seconds)."""
    desc = scenes.soup_scene(5000, 0.05)
    e = Emu(emu, desc)
    o = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(o)
    rays = scenes.random_rays(64)
    good = rays.copy()
    rays["origin"][0, 0] = np.nan; rays["direction"][1, 1] = np.nan; rays["origin"][2, 2] = np.inf; rays["direction"][3, 0] = -np.inf
    rays["direction"][4] = 0.0
    hits, occ, ctr = e.trace(rays)
    ref = o.trace_closest(rays)
    assert (hits["inst"][:5] == -1).all() and (ref["inst"][:5] == -1).all() and (occ[:5] == 0).all()
    assert np.array_equal(hits["t"][:4], rays["tmax"][:4])
    hits_good, _, ctr_good = e.trace(good)
    assert np.array_equal(hits[5:], hits_good[5:])
    assert ctr[0] <= ctr_good[0]  # the bad rays visit no nodes at all


def test_coincident_triangles_tie_break(emu):
    t = scenes.make_triangles(np.zeros((3, 3), np.float32), np.tile([[1, 0, 0]], (3, 1)).astype(np.float32), np.tile([[0, 1, 0]], (3, 1)).astype(np.float32))
    desc = scenes.SceneDesc(); desc.meshes[0] = t; desc.instances[0] = scenes.to_column_major([scenes.identity(), scenes.identity()]); desc.materials = scenes.material()
    e = Emu(emu, desc)
    rays = np.zeros(1, wire.RAY); rays["origin"] = (0.2, 0.2, 1); rays["direction"] = (0, 0, -1); rays["tmin"] = 1e-4; rays["tmax"] = 1e26
    hits, _, _ = e.trace(rays)
    assert hits["inst"][0] == 0 and hits["prim"][0] == 0


def test_authored_asset_precision(emu, oracle_mod):
    """Primary rays at a distance of ~500 triangle sizes (CesiumMan, config C1): the triangle test's t and barycentrics
    must stay within the parity tolerances there, not only on the unit-cube soups (a triangle test built on
    un-projected triple products loses (distance / size)^2 ulps and fails this)."""
    import os

    from rfw_rs_b200 import gltf

    gold_dir = os.path.join(HERE, "golden")
    asset = gltf.load_npz(os.path.join(gold_dir, "cesium_man.npz"))
    flat = gltf.flatten(asset)
    w, h = 1280, 720
    view = gltf.c1_camera(flat, w, h)
    o = oracle_mod.OracleBackend(det_eps=0.0); flat.apply(o)
    rays = np.ascontiguousarray(o.primary_rays(view, w, h).reshape(h, w)[::3, ::3].reshape(-1))
    e = Emu(emu, flat)
    hits, occ, _ = e.trace(rays)
    ref = o.trace_closest(rays, mode=oracle_mod.MODE_BVH2)
    assert (ref["inst"] >= 0).sum() > 5000
    parity.compare_hits(rays, hits, ref, parity.lookup_from_desc(flat), "cesium_man/emu", max_fraction=1e-4)
