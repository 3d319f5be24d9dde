#!/usr/bin/env python
"""Golden vectors made by the REFERENCE's own shader sources (oracle/_ref/libref_glsl.so: backends/gpu-rt/shaders/*.glsl,
*.comp compiled for the host where they lie — recipe oracle/ref_glsl/Makefile).  Run HERE, where /root/reference exists;
the GPU box and any machine without the reference read the committed tests/golden/ref_glsl_golden.npz through
tests/test_ref_golden.py (CPU tier: the oracle against it; GPU tier: the CUDA path against it).

  mt_*      intersection.glsl:1-70     8 192 (triangle, ray) pairs -> hit, t, u, v, occludes
  node_*    intersection.glsl:106-168  4 096 (MBVH node, ray) pairs -> any, per-child result, sorted tmin words
  bsdf_*    disney.glsl                4 096 (material, frame, wo, wi, r) -> eval, pdf, sampled wi, pdf, back-facing eval
  light_*   shade.comp:414-528         4 096 samples over the four light types of scenes.lights_and_lobes_scene(grid=3, subdiv=1)
  rng_*     random.glsl                wang_hash / randf sequences; bary_*: RandomBarycentrics; so_*: safe_origin
  trace_*   ray_gen.comp:202-362, ray_shadow.comp:83-243 on the oracle-built MBVH of two scenes -> closest hits, any-hit flags
  img_*     RayTracer::render host loop (lib.rs:1685-1729) running ray_gen/shade/ray_extend/ray_shadow/blit.comp: accumulators of
            three 96x54 scenes, 4 frames at sample indices 256.. (hash RNG branch), 3 segments, clamp 10
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_glsl  # noqa: E402
from rfw_rs_b200 import scenes  # noqa: E402
from tests import test_ref_glsl as T  # noqa: E402  (input generators shared with the live comparison)

HERE = os.path.dirname(os.path.abspath(__file__))
vp = T._vp


def golden_scenes():
    """name -> (scene description, camera view, w, h); shared with tests/test_ref_golden.py"""
    w, h = 96, 54
    return {
        "instanced": (scenes.instanced_scene(grid=6, subdiv=1, n_lights=4), scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h), w, h),
        "lobes": (scenes.lights_and_lobes_scene(grid=4, subdiv=2), scenes.camera_view((0.0, 3.2, -7.5), (0.0, -0.38, 1.0), w, h, aperture=0.05), w, h),
        "textured": (scenes.textured_scene(grid=3, subdiv=2, tex_size=32, skybox=True), scenes.camera_view((0.0, 2.6, -6.0), (0.0, -0.35, 1.0), w, h), w, h),
    }


def trace_cases():
    return {
        "soup": (scenes.soup_scene(30000, 0.02), scenes.random_rays(20000)),
        "instanced": (scenes.instanced_scene(grid=8, subdiv=2, n_lights=4), T._scene_rays(20000, -5.0, 5.0)),
    }


def main():
    assert ref_glsl.available(), "build oracle/_ref first: make -C oracle/ref_glsl"
    L = ref_glsl.lib()
    out = {"about": np.frombuffer(L.ref_glsl_about(), dtype=np.uint8)}
    # ---- triangle test ---------------------------------------------------------------------------------------------------
    n = 8192
    tris, rays = T._pairs(n, np.random.default_rng(101))
    hit = np.zeros(n, np.int32); tuv = np.zeros((n, 3), np.float32); occ = np.zeros(n, np.int32)
    L.ref_intersect(vp(tris), vp(rays), n, vp(hit), vp(tuv), vp(occ))
    out.update(mt_v0=tris["vertex0"], mt_v1=tris["vertex1"], mt_v2=tris["vertex2"], mt_rays=rays.view(np.float32).reshape(n, 8), mt_hit=hit, mt_tuv=tuv, mt_occ=occ)
    # ---- 4-wide node test ------------------------------------------------------------------------------------------------
    n = 4096
    rng = np.random.default_rng(102)
    lo = rng.uniform(-1.0, 1.0, size=(n, 4, 3)); ext = 10.0 ** rng.uniform(-2.0, 0.3, size=(n, 4, 3)); hi = lo + ext
    m = np.zeros((n, 32), np.float32)
    for a in range(3):
        m[:, 8 * a:8 * a + 4] = lo[:, :, a]; m[:, 8 * a + 4:8 * a + 8] = hi[:, :, a]
    nr = scenes.random_rays(n, lo=-1.5, hi=1.5)
    nr["direction"] = T._unit((lo[:, 0] + ext[:, 0] * rng.uniform(-0.3, 1.3, size=(n, 3))) - nr["origin"])
    nr["tmax"] = np.where(rng.uniform(size=n) < 0.5, 1e26, rng.uniform(0.1, 3.0, n)).astype(np.float32)
    o4 = np.zeros((n, 9), np.uint32)
    L.ref_intersect_nodes(None, vp(m), vp(nr), n, None, vp(o4))
    out.update(node_mbvh=m, node_rays=nr.view(np.float32).reshape(n, 8), node_out=o4)
    # ---- BSDF ------------------------------------------------------------------------------------------------------------
    n = 4096
    mats, args = T._bsdf_inputs(n, 103)
    bs = np.zeros((n, 12), np.float32)
    L.ref_bsdf_batch(vp(mats), n, *[vp(a) for a in args], vp(bs))
    out.update(bsdf_mats=mats.view(np.uint8).reshape(n, 96), bsdf_out=bs, **{"bsdf_" + k: a for k, a in zip(("N", "T", "B", "wo", "wi", "r"), args)})
    # ---- lights ----------------------------------------------------------------------------------------------------------
    desc = scenes.lights_and_lobes_scene(grid=3, subdiv=1)
    rb = ref_glsl.RefBackend(); desc.apply(rb)
    rng = np.random.default_rng(104)
    r0 = rng.uniform(size=n).astype(np.float32)
    I = rng.uniform(-3.0, 3.0, size=(n, 3)).astype(np.float32); I[:, 1] = rng.uniform(0.0, 1.0, n)
    Nl = T._unit(rng.normal(size=(n, 3)) + np.array([0.0, 1.5, 0.0]))
    lo_r = np.zeros((n, 8), np.float32)
    L.ref_light_batch(n, vp(r0), vp(I), vp(Nl), vp(lo_r))
    out.update(light_r0=r0, light_I=I, light_N=Nl, light_out=lo_r)
    # ---- RNG, barycentrics, safe_origin ----------------------------------------------------------------------------------
    seeds = np.concatenate([[0, 1, 12345, 0xDEADBEEF, 0xFFFFFFFF], rng.integers(0, 2**32, 507)]).astype(np.uint32)
    hashes = np.array([L.ref_wang_hash(int(s)) for s in seeds], np.uint32)
    seq = np.zeros((len(seeds), 4), np.float32)
    for k, s in enumerate(seeds):
        st = C.c_uint32(int(s) | 1)
        for j in range(4):
            seq[k, j] = L.ref_randf(C.byref(st))
    xs = np.concatenate([rng.uniform(size=1019), [0.0, 0.25, 0.5, 0.75, 0.999999]]).astype(np.float32)
    bary = np.zeros((len(xs), 3), np.float32)
    for k, x in enumerate(xs):
        L.ref_random_barycentrics(C.c_float(float(x)), vp(bary[k]))
    so_in = np.zeros((1024, 9), np.float32); so_out = np.zeros((1024, 3), np.float32)
    for k in range(1024):
        so_in[k, :3] = rng.normal(size=3) * 10.0 ** rng.integers(-3, 3); so_in[k, 3:6] = T._unit(rng.normal(size=(1, 3)))[0]; so_in[k, 6:] = T._unit(rng.normal(size=(1, 3)))[0]
        L.ref_safe_origin(vp(so_in[k, :3]), vp(so_in[k, 3:6]), vp(so_in[k, 6:]), vp(so_out[k]))
    out.update(rng_seeds=seeds, rng_hash=hashes, rng_seq=seq, bary_in=xs, bary_out=bary, so_in=so_in, so_out=so_out)
    # ---- traversal -------------------------------------------------------------------------------------------------------
    for name, (desc, rays) in trace_cases().items():
        rb = ref_glsl.RefBackend(); desc.apply(rb)
        hits = rb.trace_closest(rays)
        out["trace_" + name + "_hits"] = hits.view(np.uint8).reshape(len(hits), 20)
        out["trace_" + name + "_occ"] = rb.trace_any(rays).astype(np.uint8)
    # ---- frames ----------------------------------------------------------------------------------------------------------
    for name, (desc, view, w, h) in golden_scenes().items():
        rb = ref_glsl.RefBackend(); desc.apply(rb)
        acc, img, ctr = rb.render(view, w, h, 4, depth=3, first_sample=256)
        out["img_" + name + "_acc"] = acc
        out["img_" + name + "_counts"] = np.array([ctr["extension_rays"], ctr["shadow_rays"]], np.uint64)
    path = os.path.join(HERE, "ref_glsl_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
