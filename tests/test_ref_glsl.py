"""The oracle (and the product's host-compiled bodies) held against the REFERENCE's own shader sources.

oracle/_ref/libref_glsl.so is backends/gpu-rt/shaders/{intersection,disney,lambert,utils,random,structs}.glsl and the five
compute kernels {ray_gen,ray_extend,shade,ray_shadow,blit}.comp compiled for the host where they lie, against the glm the
reference vendors (recipe: oracle/ref_glsl/Makefile).  That is "outputs of the reference itself run here" — what pins the
oracle.  The library exists wherever /root/reference does (this container); elsewhere the tests skip and the
committed fixtures made from it (tests/golden/ref_glsl_golden.npz, test_ref_golden.py) take over.

Function level: bit-equality wherever the arithmetic is +, -, *, /, sqrt (IEEE, same operation order); a few ulp where libm
transcendentals (sin, cos, log, exp) are involved.  Kernel level: identical ray counts and images equal to rounding."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from rfw_rs_b200 import scenes, wire

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def ref(oracle_mod):
    from oracle import ref_glsl

    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_glsl"), "-s"])
    if not ref_glsl.available():
        pytest.skip("oracle/_ref/libref_glsl.so not built (/root/reference absent)")
    return ref_glsl


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def _unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _ulps(a, b):
    """distance in float32 representable values (same-sign finite inputs)"""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _pairs(n, rng):
    """n (triangle, ray) pairs: rays aimed at points inside, on the edges of, and just outside their triangle, from both sides,
    plus unrelated pairs and limit cases (t near tmin / tmax)."""
    c = rng.uniform(-2.0, 2.0, size=(n, 3))
    scale = 10.0 ** rng.uniform(-2.0, 0.5, size=(n, 1))
    v0, v1, v2 = (c + rng.normal(size=(n, 3)) * scale for _ in range(3))
    tris = scenes.make_triangles(v0.astype(np.float32), v1.astype(np.float32), v2.astype(np.float32))
    b = rng.uniform(-0.15, 1.15, size=(n, 2))
    kind = rng.integers(0, 8, n)
    b[kind == 0, 0] = 0.0                      # on the edge u = 0
    b[kind == 1, 1] = 0.0                      # on the edge v = 0
    b[kind == 2, 1] = 1.0 - b[kind == 2, 0]    # on the edge u + v = 1
    target = v0 + (v1 - v0) * b[:, :1] + (v2 - v0) * b[:, 1:]
    target[kind == 3] = rng.uniform(-2.0, 2.0, size=((kind == 3).sum(), 3))  # unrelated
    org = target + _unit(rng.normal(size=(n, 3))) * 10.0 ** rng.uniform(-1.5, 1.0, size=(n, 1))
    d = target - org
    dist = np.linalg.norm(d, axis=1)
    d = d / dist[:, None]
    not_norm = kind == 4
    d[not_norm] *= rng.uniform(0.2, 3.0, size=(not_norm.sum(), 1))           # un-normalised directions (object-space rays)
    rays = np.zeros(n, wire.RAY)
    rays["origin"] = org.astype(np.float32); rays["direction"] = d.astype(np.float32)
    rays["tmin"] = 1e-4; rays["tmax"] = 1e26
    lim = kind == 5
    rays["tmax"][lim] = (dist[lim] * rng.choice([0.5, 0.999999, 1.0, 1.000001, 2.0], lim.sum())).astype(np.float32)
    lim = kind == 6
    rays["tmin"][lim] = (dist[lim] * rng.choice([0.5, 0.999999, 1.0, 1.000001, 2.0], lim.sum())).astype(np.float32)
    return np.ascontiguousarray(tris), rays


def test_triangle_test_is_the_references_bit_for_bit(ref, oracle_mod):
    """intersection.glsl:1-38 (closest) and :40-70 (any-hit) against the oracle's mt_intersect on 10^6 pairs: same accept /
    reject decision and bit-identical t, u, v.  det_eps = 1e-4 is the reference's own determinant epsilon."""
    n = 1_000_000
    tris, rays = _pairs(n, np.random.default_rng(11))
    L, O = ref.lib(), oracle_mod.lib()
    hit_r = np.zeros(n, np.int32); tuv_r = np.zeros((n, 3), np.float32); occ_r = np.zeros(n, np.int32)
    hit_o = np.zeros(n, np.int32); tuv_o = np.zeros((n, 3), np.float32); occ_o = np.zeros(n, np.int32)
    L.ref_intersect(_vp(tris), _vp(rays), n, _vp(hit_r), _vp(tuv_r), _vp(occ_r))
    O.orc_triangle_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    O.orc_triangle_batch(_vp(tris), _vp(rays), n, 1e-4, _vp(hit_o), _vp(tuv_o), _vp(occ_o))
    assert 0.15 < hit_r.mean() < 0.8                      # both outcomes are well represented (small triangles fall to the 1e-4 determinant epsilon)
    assert np.array_equal(hit_r, hit_o)
    assert np.array_equal(occ_r, occ_o) and np.array_equal(occ_r, hit_r)
    h = hit_r == 1
    assert np.array_equal(tuv_r[h].view(np.uint32), tuv_o[h].view(np.uint32))
    assert np.array_equal(tuv_r[~h, 0], rays["tmax"][~h])  # t untouched on a miss (inout)
    # det_eps = 0 (what the soups of the parity runs use: DESIGN.md §2) only ADDS hits the reference's epsilon rejects
    O.orc_triangle_batch(_vp(tris), _vp(rays), n, 0.0, _vp(hit_o), _vp(tuv_o), None)
    assert (hit_o >= hit_r).all() and np.array_equal(tuv_r[h].view(np.uint32), tuv_o[h].view(np.uint32))


def test_node_tests_match_the_reference(ref, oracle_mod):
    """intersect_node (intersection.glsl:72-92) and intersect_mnode (:106-168): same per-child decisions and the same
    near-to-far order (the sorted tmin words carry the child index in their two mantissa LSBs) on finite, non-degenerate
    inputs.  Where the oracle deviates on purpose (`tmin <= t` instead of `<`: canonical ties) the inputs
    have no exact ties, so the decisions must coincide."""
    n = 400_000
    rng = np.random.default_rng(12)
    lo = rng.uniform(-1.0, 1.0, size=(n, 4, 3)); ext = 10.0 ** rng.uniform(-2.0, 0.3, size=(n, 4, 3))
    hi = lo + ext
    m = np.zeros(n, dtype=np.dtype([("min_x", np.float32, 4), ("max_x", np.float32, 4), ("min_y", np.float32, 4), ("max_y", np.float32, 4), ("min_z", np.float32, 4),
                                    ("max_z", np.float32, 4), ("children", np.int32, 4), ("counts", np.int32, 4)]))
    for a, nm in enumerate("xyz"):
        m["min_" + nm] = lo[:, :, a]; m["max_" + nm] = hi[:, :, a]
    empty = rng.uniform(size=(n, 4)) < 0.15          # empty child slots: inverted boxes, as the builders write them
    for nm in "xyz":
        m["min_" + nm][empty] = 1e34; m["max_" + nm][empty] = -1e34
    b2 = np.zeros(n, dtype=np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("left_first", np.int32), ("count", np.int32)]))
    b2["bmin"] = lo[:, 0]; b2["bmax"] = hi[:, 0]
    rays = scenes.random_rays(n, lo=-1.5, hi=1.5)
    aim = (lo[:, 0] + ext[:, 0] * rng.uniform(-0.3, 1.3, size=(n, 3))) - rays["origin"]   # most rays pass near child 0 (= the BVH2 box)
    rays["direction"] = _unit(aim) * np.where(rng.uniform(size=(n, 1)) < 0.5, 1.0, -1.0).astype(np.float32)
    rays["tmax"] = np.where(rng.uniform(size=n) < 0.5, 1e26, rng.uniform(0.1, 3.0, n)).astype(np.float32)
    L, O = ref.lib(), oracle_mod.lib()
    o2r = np.zeros((n, 3), np.float32); o4r = np.zeros((n, 9), np.uint32); o2o = np.zeros((n, 3), np.float32); o4o = np.zeros((n, 9), np.uint32)
    L.ref_intersect_nodes(_vp(b2), _vp(m), _vp(rays), n, _vp(o2r), _vp(o4r))
    O.orc_node_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    O.orc_node_batch(_vp(b2), _vp(m), _vp(rays), n, _vp(o2o), _vp(o4o))
    assert 0.05 < o2r[:, 0].mean() < 0.9 and 0.2 < o4r[:, 0].mean() < 0.99
    assert np.array_equal(o2r[:, 0], o2o[:, 0])
    assert np.array_equal(o2r[:, 1].view(np.uint32), o2o[:, 1].view(np.uint32))      # entry distance, bit for bit
    assert np.array_equal(o4r[:, :5], o4o[:, :5])                                     # any + the four per-child decisions
    anyhit = o4r[:, 0] == 1
    assert np.array_equal(o4r[anyhit, 5:], o4o[anyhit, 5:])                           # sorted entry distances incl. the index bits


def _random_materials(n, rng):
    mats = np.concatenate([scenes.material(color=rng.uniform(0.05, 1.0, 3), metallic=rng.choice([0.0, 1.0, rng.uniform()]), roughness=rng.uniform(0.02, 1.0),
                                           specular_f=rng.uniform(), subsurface=rng.choice([0.0, rng.uniform()]), specular=rng.uniform(0.2, 1.0, 3),
                                           transmission=rng.choice([0.0, 1.0, rng.uniform()]), eta=rng.uniform(0.4, 1.0), clearcoat=rng.choice([0.0, rng.uniform()]),
                                           clearcoat_gloss=rng.uniform(), specular_tint=rng.uniform(), absorption=rng.uniform(0.0, 1.0, 3)) for _ in range(256)])
    return np.ascontiguousarray(mats[rng.integers(0, len(mats), n)])


def _bsdf_inputs(n, seed):
    rng = np.random.default_rng(seed)
    mats = _random_materials(n, rng)
    N = _unit(rng.normal(size=(n, 3)))
    T = _unit(np.cross(N, _unit(rng.normal(size=(n, 3)))))
    B = np.cross(N, T).astype(np.float32)
    wo = _unit(rng.normal(size=(n, 3))); wo = np.where((np.einsum("ij,ij->i", wo, N) < 0)[:, None], -wo, wo).astype(np.float32)
    wi = _unit(rng.normal(size=(n, 3)))
    r = rng.uniform(size=(n, 2)).astype(np.float32)
    return mats, [np.ascontiguousarray(a) for a in (N, T, B, wo, wi, r)]


def test_disney_bsdf_matches_the_reference(ref, oracle_mod):
    """disney.glsl: BSDFEval / BSDFPdf / BSDFSample (through the calls shade.comp makes) on 10^5 random (material, frame, wo, wi,
    r3, r4) with every lobe active somewhere: the oracle restates the same operations in the same order, so values are equal
    up to the libm calls inside (sin / cos / log / exp); the scalar building blocks GTR1/GTR2/SmithGGX/Fr/Schlick are bit-equal
    except GTR1 (log)."""
    n = 100_000
    mats, args = _bsdf_inputs(n, 13)
    L, O = ref.lib(), oracle_mod.lib()
    out_r = np.zeros((n, 12), np.float32); out_o = np.zeros((n, 12), np.float32)
    L.ref_bsdf_batch(_vp(mats), n, *[_vp(a) for a in args], _vp(out_r))
    O.orc_bsdf_batch(_vp(mats), C.c_uint32(n), *[_vp(a) for a in args], _vp(out_o))
    assert np.isfinite(out_r[:, :4]).all() and (out_r[:, :3].max(axis=1) > 0).mean() > 0.5
    for name, cols in (("eval", slice(0, 3)), ("pdf", slice(3, 4)), ("sampled direction", slice(4, 7)), ("sample pdf", slice(7, 8)), ("eval back-facing", slice(8, 11))):
        a, b = out_r[:, cols], out_o[:, cols]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        close = same | (_ulps(np.nan_to_num(a), np.nan_to_num(b)) <= 4)
        assert same.mean() > 0.999, (name, float(same.mean()))
        assert close.all(), (name, a[~close][:4], b[~close][:4])


def test_lambert_and_utils_match_the_reference(ref, oracle_mod, shade_emu_lib):
    """utils.glsl (safe_origin, CLAMPINTENSITY, tangent space, normal packing), random.glsl (wang_hash, xorshift, randf),
    shade.comp:372-412 (RandomBarycentrics): bit-equal between the reference, the oracle and the PRODUCT's shading.cuh
    compiled for the host (tests/hostemu/shade_emu.cpp)."""
    L, O, E = ref.lib(), oracle_mod.lib(), shade_emu_lib
    rng = np.random.default_rng(14)
    O.orc_randf.restype = C.c_float
    for s in [0, 1, 12345, 0xDEADBEEF, 0xFFFFFFFF] + [int(x) for x in rng.integers(0, 2**32, 2000)]:
        assert L.ref_wang_hash(s) == O.orc_wang_hash(C.c_uint32(s)) == E.emu_wang_hash(s)
        sr, so, se = C.c_uint32(s | 1), C.c_uint32(s | 1), C.c_uint32(s | 1)
        for _ in range(4):
            a, b, c = L.ref_randf(C.byref(sr)), O.orc_randf(C.byref(so)), E.emu_randf(C.addressof(se))
            assert a == b == c and sr.value == so.value == se.value
    br, bo, be = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    for x in np.concatenate([rng.uniform(size=4000), [0.0, 0.25, 0.5, 0.75, 0.999999]]).astype(np.float32):
        L.ref_random_barycentrics(C.c_float(float(x)), _vp(br)); O.orc_random_barycentrics(C.c_float(float(x)), _vp(bo)); E.emu_random_barycentrics(float(x), be.ctypes.data)
        assert np.array_equal(br, bo), (x, br, bo)
        assert np.allclose(br, be, rtol=0, atol=2e-7)   # the product's branch-free form sums the same halvings in another order
    for k in range(4000):
        O3 = (rng.normal(size=3) * 10.0 ** rng.integers(-3, 3)).astype(np.float32); R3 = _unit(rng.normal(size=(1, 3)))[0]; N3 = _unit(rng.normal(size=(1, 3)))[0]
        L.ref_safe_origin(_vp(O3), _vp(R3), _vp(N3), _vp(br)); O.orc_safe_origin(_vp(O3), _vp(R3), _vp(N3), _vp(bo)); E.emu_safe_origin(O3.ctypes.data, R3.ctypes.data, N3.ctypes.data, be.ctypes.data)
        assert np.array_equal(br, bo) and np.array_equal(br, be), (O3, R3, N3, br, bo, be)


@pytest.fixture(scope="module")
def shade_emu_lib():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostemu"), "-s"])
    L = C.CDLL(os.path.join(HERE, "hostemu", "libshade_emu.so"))
    vp = C.c_void_p
    L.emu_bsdf_batch.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, vp, vp]
    L.emu_light_batch.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.emu_wang_hash.argtypes = [C.c_uint32]; L.emu_wang_hash.restype = C.c_uint32
    L.emu_randf.argtypes = [vp]; L.emu_randf.restype = C.c_float
    L.emu_random_barycentrics.argtypes = [C.c_float, vp]
    L.emu_safe_origin.argtypes = [vp, vp, vp, vp]
    return L


def _close(a, b, rtol, atol):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))


def test_product_shading_source_matches_the_reference(ref, shade_emu_lib):
    """The PRODUCT's shading.cuh (compiled for the host, unmodified) directly against the reference's disney.glsl and
    shade.comp light sampling — no oracle in between.  The product source is compiled with FMA contraction and restructured
    (SoA loads, branch-free barycentrics), so agreement is to rounding; disagreements beyond it sit on branch thresholds."""
    n = 60_000
    mats, args = _bsdf_inputs(n, 15)
    L, E = ref.lib(), shade_emu_lib
    out_r = np.zeros((n, 12), np.float32); out_e = np.zeros((n, 12), np.float32)
    L.ref_bsdf_batch(_vp(mats), n, *[_vp(a) for a in args], _vp(out_r))
    E.emu_bsdf_batch(mats.ctypes.data, n, *[a.ctypes.data for a in args], out_e.ctypes.data)
    for name, cols in (("eval", slice(0, 3)), ("pdf", slice(3, 4)), ("sampled direction", slice(4, 7)), ("sample pdf", slice(7, 8)), ("eval back-facing", slice(8, 11))):
        a, b = out_e[:, cols], out_r[:, cols]
        ok = (_close(a, b, 2e-4, 1e-6) | (np.isnan(a) & np.isnan(b))).all(axis=1)
        assert ok.mean() > (0.995 if "pdf" in name else 0.998), (name, float(ok.mean()), a[~ok][:3], b[~ok][:3])  # GGX pdfs amplify a last-bit difference of N.H
    desc = scenes.lights_and_lobes_scene(grid=3, subdiv=1)
    rb = ref.RefBackend(); desc.apply(rb)
    rng = np.random.default_rng(16)
    r0 = rng.uniform(size=n).astype(np.float32)
    I = rng.uniform(-3.0, 3.0, size=(n, 3)).astype(np.float32); I[:, 1] = rng.uniform(0.0, 1.0, n)
    Nl = _unit(rng.normal(size=(n, 3)) + np.array([0.0, 1.5, 0.0]))
    lo_r = np.zeros((n, 8), np.float32); lo_e = np.zeros((n, 8), np.float32)
    al, pl, sl, dl = (np.ascontiguousarray(x) for x in (desc.area_lights, desc.point_lights, desc.spot_lights, desc.directional_lights))
    L.ref_light_batch(n, _vp(r0), _vp(I), _vp(Nl), _vp(lo_r))
    E.emu_light_batch(al.ctypes.data, len(al), pl.ctypes.data, len(pl), sl.ctypes.data, len(sl), dl.ctypes.data, len(dl), n, r0.ctypes.data, I.ctypes.data, Nl.ctypes.data,
                      lo_e.ctypes.data)
    assert (lo_r[:, 4] > 0).mean() > 0.3 and len(np.unique(np.round(lo_r[:, 5:8], 3), axis=0)) >= 4
    ok = _close(lo_e, lo_r, 2e-4, 1e-5).all(axis=1)
    assert ok.mean() > 0.999, (float(ok.mean()), lo_e[~ok][:3], lo_r[~ok][:3])


def test_light_sampling_matches_the_reference(ref, oracle_mod):
    """RandomPointOnLight (shade.comp:414-528) for all four light types with the lights bound as the reference binds them:
    sampled point, pick probability, pdf and radiance bit-equal to the oracle's restatement."""
    n = 100_000
    desc = scenes.lights_and_lobes_scene(grid=3, subdiv=1)
    rb = ref.RefBackend(); desc.apply(rb)
    rng = np.random.default_rng(17)
    r0 = rng.uniform(size=n).astype(np.float32)
    I = rng.uniform(-3.0, 3.0, size=(n, 3)).astype(np.float32); I[:, 1] = rng.uniform(0.0, 1.0, n)
    Nl = _unit(rng.normal(size=(n, 3)) + np.array([0.0, 1.5, 0.0]))
    lo_r = np.zeros((n, 8), np.float32); lo_o = np.zeros((n, 8), np.float32)
    ref.lib().ref_light_batch(n, _vp(r0), _vp(I), _vp(Nl), _vp(lo_r))
    oracle_mod.lib().orc_light_batch(rb.o.h, C.c_uint32(n), _vp(r0), _vp(I), _vp(Nl), _vp(lo_o))
    assert (lo_r[:, 4] > 0).mean() > 0.3 and len(np.unique(np.round(lo_r[:, 5:8], 3), axis=0)) >= 4
    assert np.array_equal(lo_r.view(np.uint32), lo_o.view(np.uint32))


def test_eye_rays_match_the_reference(ref, oracle_mod):
    """generate_eye_ray (ray_gen.comp:103-146; thin lens, 9-blade aperture, hash RNG branch) and the pinhole generate_ray
    (:93-101 = CameraView3D::generate_ray, structs.rs:549-556)."""
    w, h = 160, 90
    L, O = ref.lib(), oracle_mod.lib()
    for aperture in (1e-4, 0.08):
        view = scenes.camera_view((0.3, 2.0, -5.0), (0.1, -0.3, 1.0), w, h, aperture=aperture)
        v = np.ascontiguousarray(view)
        out_o = np.zeros((w * h, 6), np.float32)
        O.orc_eye_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        O.orc_eye_rays(_vp(v), w, h, w * h, 300, _vp(out_o))
        L.ref_set_camera(_vp(v), w, h, 300)
        od = np.zeros(6, np.float32)
        worst = 0
        for p in range(0, w * h, 7):
            seed = O.orc_wang_hash(C.c_uint32((p * 16789 + 300 * 1791) & 0xFFFFFFFF))
            L.ref_eye_ray(p, seed, _vp(od))
            worst = max(worst, int(_ulps(od, out_o[p]).max()))
        assert worst <= 2, worst   # cos / sin of the blade angle go through libm on both sides: equal up to the last bit
        rays = oracle_mod.OracleBackend().primary_rays(view, w, h)
        for p in range(0, w * h, 11):
            L.ref_pinhole_ray(p, _vp(od))
            assert np.array_equal(od[:3], rays["origin"][p]) and np.array_equal(od[3:], rays["direction"][p])


def _scene_rays(n, lo, hi, seed=5678):
    rays = scenes.random_rays(n, lo=lo, hi=hi, seed=seed)
    rays["origin"][:, 1] = np.abs(rays["origin"][:, 1]) * 0.5 + 0.05
    return rays


@pytest.mark.parametrize("which", ["soup", "instanced", "lobes"])
def test_traversal_loops_match_the_reference(ref, oracle_mod, which):
    """The reference's own traversal loops (intersect_top_mbvh / intersect_mbvh of ray_gen.comp:202-250,310-362; any-hit forms
    of ray_shadow.comp:83-132,191-243) walking the trees the oracle built: the oracle's MODE_MBVH traversal must return the same
    (instance, primitive) and bit-identical t, u, v.  The only sanctioned difference is an EXACT t tie between two triangles,
    where the reference keeps the first one visited and the oracle the canonical one."""
    if which == "soup":
        desc, rays = scenes.soup_scene(30000, 0.02), scenes.random_rays(60000)
    elif which == "instanced":
        desc, rays = scenes.instanced_scene(grid=8, subdiv=2, n_lights=4), _scene_rays(60000, -5.0, 5.0)
    else:
        desc, rays = scenes.lights_and_lobes_scene(grid=4, subdiv=2), _scene_rays(60000, -4.0, 4.0)
    rb = ref.RefBackend(); desc.apply(rb)
    o = oracle_mod.OracleBackend(det_eps=1e-4); desc.apply(o)
    h_ref = rb.trace_closest(rays)
    h_orc = o.trace_closest(rays, mode=oracle_mod.MODE_MBVH)
    assert (h_ref["inst"] >= 0).mean() > 0.15
    same_id = (h_ref["inst"] == h_orc["inst"]) & (h_ref["prim"] == h_orc["prim"])
    # (object-space origins are M^-1 * (O, 1): the oracle sums in glm's order, oracle/vecmath.h::xform_point)
    assert same_id.all(), (~same_id).sum()
    for f in ("t", "u", "v"):
        assert np.array_equal(h_ref[f].view(np.uint32), h_orc[f].view(np.uint32)), f
    occ_ref = rb.trace_any(rays)
    occ_orc = o.trace_any(rays, mode=oracle_mod.MODE_MBVH)
    assert np.array_equal(occ_ref, occ_orc)
    assert np.array_equal(occ_ref == 1, h_ref["inst"] >= 0)
    # the BVH2 loops of the reference: dead code behind `USE_MBVH 1`, with the _ltmin defect (visiting order only) and a
    # strict `t_max > t_min` that rejects every flat box (ground quads) — comparable on the soup only, whose boxes have volume
    if which == "soup":
        h_ref2 = rb.trace_closest(rays[:20000], mode=1)
        assert ((h_ref2["inst"] == h_ref["inst"][:20000]) & (h_ref2["prim"] == h_ref["prim"][:20000])).mean() > 0.9999


@pytest.mark.parametrize("which", ["instanced", "lobes", "textured"])
def test_rendered_image_matches_the_reference_kernels(ref, oracle_mod, which):
    """Whole frames: ray_gen -> shade -> (ray_extend -> shade)* -> ray_shadow -> blit run by the host loop of
    RayTracer::render (lib.rs:1685-1729), i.e. the reference renderer itself on the CPU, against the oracle's per-path
    restatement with the same settings (3 segments, clamp 10, sample indices >= 256 = hash RNG branch).  Same number of
    extension and shadow rays; radiance equal to rounding (operation order inside glm's mat4 * vec4 and the accumulation
    order differ, and GGX with small roughness amplifies a last-bit difference of N.H)."""
    w, h, spp, depth = 96, 54, 4, 3
    sky = (0.0, 0.0, 0.0)
    if which == "instanced":
        desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
        view = scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h)
    elif which == "lobes":
        desc = scenes.lights_and_lobes_scene(grid=4, subdiv=2)
        view = scenes.camera_view((0.0, 3.2, -7.5), (0.0, -0.38, 1.0), w, h, aperture=0.05)
    else:
        desc = scenes.textured_scene(grid=3, subdiv=2, tex_size=32, skybox=True)
        view = scenes.camera_view((0.0, 2.6, -6.0), (0.0, -0.35, 1.0), w, h)
    rb = ref.RefBackend(); desc.apply(rb)
    o = oracle_mod.OracleBackend(det_eps=1e-4); desc.apply(o)
    acc_r, img_r, ctr = rb.render(view, w, h, spp, depth=depth, first_sample=256)
    acc_o, st = o.render(view, w, h, spp, depth, sky=sky, first_sample=256)
    assert ctr["extension_rays"] == st["extension_rays"] and ctr["shadow_rays"] == st["shadow_rays"], (ctr, st)
    assert st["shadow_rays"] > 1000
    # a miss whose |D.y| exceeds 1 by an ulp makes the reference's acos (shade.comp:91) return NaN (sampled directions are not unit length: the sampling frame is built from gN with the shading normal's
    # tangents, disney.glsl:275-285): the texture coordinate and
    # with it the whole pixel is undefined there; the oracle clamps the argument.  Such pixels are left out.
    bad = ~np.isfinite(acc_r[..., :3]).all(axis=2)
    assert bad.mean() <= 5e-3 and np.isfinite(acc_o).all()
    acc_r = np.where(bad[..., None], acc_o, acc_r)
    d = (acc_r[..., :3] - acc_o[..., :3]).astype(np.float64) / spp
    rel = np.abs(d) / np.maximum(1e-2, np.abs(acc_o[..., :3]) / spp)
    assert float(np.sqrt(np.mean(d ** 2))) <= 2e-5, float(np.sqrt(np.mean(d ** 2)))
    assert (rel.max(axis=2) <= 1e-5).mean() > 0.97
    assert rel.max() <= 2e-2
    # blit.comp: sqrt(acc / (sample_count + 1)) with the sample index of the last frame
    assert np.allclose(img_r[..., :3][~bad], np.sqrt(acc_r[..., :3] / (256 + spp))[~bad], rtol=1e-6, atol=1e-7)


# ---- the blue-noise sampler of the first 256 samples -------------------------------------------------------------------------
def reference_blue_noise_table():
    """The u32 buffer create_blue_noise_buffer() builds (backends/gpu-rt/src/blue_noise.rs:40970-41004) from the three u64 tables of
    that file, reproduced from the DATA where it lies: each table is viewed as bytes, and only the first len * size_of::<u32>()
    bytes of it are used (the function's own slice length), repeated to fill the region.  None when /root/reference is absent."""
    path = "/root/reference/backends/gpu-rt/src/blue_noise.rs"
    if not os.path.exists(path):
        return None
    import re

    text = open(path).read()
    tables = {}
    for name, n in (("SOB256_64", 8192), ("SCR256_64", 16384), ("RNK256_64", 16384)):
        body = text[text.index(f"static {name}: [u64; {n}] = ["):]
        body = body[body.index("= [") + 3:body.index("];")]
        vals = np.array([int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", body)], dtype=np.uint64)
        assert len(vals) == n, (name, len(vals))
        tables[name] = vals.view(np.uint8)[: n * 4]   # from_raw_parts(ptr as *const u8, len * size_of::<u32>())
    buf = np.zeros(65536 * 5, np.uint32)
    buf[:65536] = tables["SOB256_64"][np.arange(65536) % len(tables["SOB256_64"])]
    k = np.arange(128 * 128 * 8)
    buf[65536:65536 + len(k)] = tables["SCR256_64"][k % len(tables["SCR256_64"])]
    buf[3 * 65536:3 * 65536 + len(k)] = tables["RNK256_64"][k % len(tables["RNK256_64"])]
    return buf


def synthetic_blue_noise_table(seed=77):
    """A table of the same shape with seeded random bytes: what the GPU-tier tests use (the real table is reference data and
    is not copied into the repository); the sampler's index arithmetic is exercised all the same."""
    return np.random.default_rng(seed).integers(0, 256, 65536 * 5).astype(np.uint32)


@pytest.mark.parametrize("table", ["reference", "synthetic"])
def test_blue_noise_sampler_matches_the_reference(ref, oracle_mod, shade_emu_lib, table):
    """blueNoiseSampler (ray_gen.comp:72-91) for every dimension a depth-5 path uses, all 128 x 128 pixels' corners and random
    ones, sample counts 0..255: reference shader == oracle == the product's shade_path.cuh body, bit for bit."""
    bn = reference_blue_noise_table() if table == "reference" else synthetic_blue_noise_table()
    if bn is None:
        pytest.skip("/root/reference absent")
    L, O, E = ref.lib(), oracle_mod.lib(), shade_emu_lib
    o = oracle_mod.OracleBackend(); o.set_blue_noise(bn)
    bn_i32 = np.ascontiguousarray(bn.astype(np.int32))  # (kept alive across the call: _vp only carries the address)
    L.ref_set_blue_noise(_vp(bn_i32), len(bn_i32))
    E.emu_set_blue_noise.argtypes = [C.c_void_p, C.c_uint32]
    E.emu_blue_noise_sample.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]; E.emu_blue_noise_sample.restype = C.c_float
    O.orc_blue_noise_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32]; O.orc_blue_noise_sample.restype = C.c_float
    keep = np.ascontiguousarray(bn)
    E.emu_set_blue_noise(_vp(keep), len(keep))
    rng = np.random.default_rng(5)
    pts = [(0, 0), (127, 127), (127, 0), (0, 127), (128, 129), (1919, 1079)] + [tuple(int(v) for v in rng.integers(0, 4096, 2)) for _ in range(300)]
    seen = set()
    for (x, y) in pts:
        for dim in list(range(24)) + [255, 256, 300]:
            for sc in (0, 1, 17, 128, 254, 255):
                a = L.ref_blue_noise_sample(x, y, dim, sc)
                b = O.orc_blue_noise_sample(o.h, x, y, dim, sc)
                c = E.emu_blue_noise_sample(x, y, dim, sc)
                assert a == b == c, (x, y, dim, sc, a, b, c)
                seen.add(a)
    assert len(seen) > 200 and all(0.0 < v < 1.0 for v in seen)
    E.emu_set_blue_noise(None, 0)


@pytest.mark.parametrize("table", ["reference", "synthetic"])
def test_blue_noise_frames_match_the_reference_kernels(ref, oracle_mod, table):
    """Frames 0..3 (sample_count < 256: the blue-noise branch of ray_gen.comp:109-115 and shade.comp:190-196,216-222) rendered by
    the reference's kernels against the oracle with the same tables — and a frame straddling sample 256, where both switch to
    the hash RNG."""
    bn = reference_blue_noise_table() if table == "reference" else synthetic_blue_noise_table()
    if bn is None:
        pytest.skip("/root/reference absent")
    w, h, depth = 96, 54, 3
    desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
    view = scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h)
    rb = ref.RefBackend(); desc.apply(rb); rb.set_blue_noise(bn.astype(np.int32))
    o = oracle_mod.OracleBackend(det_eps=1e-4); desc.apply(o); o.set_blue_noise(bn)
    hashed, _ = oracle_mod.OracleBackend(det_eps=1e-4), None
    for first, spp in ((0, 4), (254, 4)):
        acc_r, _, ctr = rb.render(view, w, h, spp, depth=depth, first_sample=first, acc=np.zeros((h, w, 4), np.float32))
        acc_o, st = o.render(view, w, h, spp, depth, sky=(0.0, 0.0, 0.0), first_sample=first)
        assert ctr["extension_rays"] == st["extension_rays"] and ctr["shadow_rays"] == st["shadow_rays"], (first, ctr, st)
        d = (acc_r[..., :3] - acc_o[..., :3]).astype(np.float64) / spp
        assert float(np.sqrt(np.mean(d ** 2))) <= 2e-5, (first, float(np.sqrt(np.mean(d ** 2))))
    # and the tables do change the image (the test would pass vacuously if neither side used them)
    o2 = oracle_mod.OracleBackend(det_eps=1e-4); desc.apply(o2)
    acc_h, _ = o2.render(view, w, h, 4, depth, sky=(0.0, 0.0, 0.0), first_sample=0)
    acc_b, _ = o.render(view, w, h, 4, depth, sky=(0.0, 0.0, 0.0), first_sample=0)
    assert float(np.abs(acc_h - acc_b).mean()) > 1e-2
