"""ctypes front-end of oracle/_ref/libref_glsl.so — the REFERENCE's own shader sources compiled for the host.

TEST INFRASTRUCTURE ONLY.  The library is built by oracle/ref_glsl/Makefile from the sources where they lie under
/root/reference (device functions of intersection / disney / lambert / utils / random.glsl and the five compute kernels
ray_gen / ray_extend / shade / ray_shadow / blit.comp) against the glm the reference vendors.  It exists in the build
container; on the GPU box the prebuilt .so travels with the snapshot.  `available()` says whether it can be loaded.

`RefBackend` replays a scene description like `OracleBackend` does, but renders and traces with the reference kernels:
the acceleration structure comes from the oracle's builder in the reference's buffer layout (rtbvh, the reference's
builder, is an un-vendored crate), everything that walks or shades it is the reference's code.
"""
import ctypes as C
import os

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_glsl.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        L.ref_glsl_about.restype = C.c_char_p
        L.ref_bind_scene.argtypes = [vp] * 8
        L.ref_bind_materials.argtypes = [vp]
        L.ref_bind_lights.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32]
        L.ref_set_texture_callback.argtypes = [vp, vp]
        L.ref_set_blue_noise.argtypes = [vp, u64]
        L.ref_render.argtypes = [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp]
        L.ref_trace_closest.argtypes = [vp, u64, vp, i32]
        L.ref_trace_any.argtypes = [vp, u64, vp, i32]
        L.ref_intersect.argtypes = [vp, vp, u64, vp, vp, vp]
        L.ref_intersect_nodes.argtypes = [vp, vp, vp, u64, vp, vp]
        L.ref_bsdf_batch.argtypes = [vp, u32] + [vp] * 7
        L.ref_lambert_batch.argtypes = [vp, u32] + [vp] * 7
        L.ref_disney_scalars.argtypes = [vp, vp, u32, vp]
        L.ref_light_batch.argtypes = [u32, vp, vp, vp, vp]
        L.ref_wang_hash.argtypes = [u32]
        L.ref_wang_hash.restype = u32
        L.ref_randf.argtypes = [vp]
        L.ref_randf.restype = f32
        L.ref_random_barycentrics.argtypes = [f32, vp]
        L.ref_safe_origin.argtypes = [vp] * 4
        L.ref_tangent_space.argtypes = [vp, vp]
        L.ref_pack_normal.argtypes = [vp]
        L.ref_pack_normal.restype = u32
        L.ref_unpack_normal.argtypes = [u32, vp]
        L.ref_clamp_intensity.argtypes = [vp, f32]
        L.ref_set_camera.argtypes = [vp, i32, i32, i32]
        L.ref_eye_ray.argtypes = [u32, u32, vp]
        L.ref_pinhole_ray.argtypes = [u32, vp]
        L.ref_blue_noise_sample.argtypes = [i32, i32, i32, i32]
        L.ref_blue_noise_sample.restype = f32
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Flat(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("triangles", "prim_indices", "bvh_nodes", "mbvh_nodes", "instances", "top_bvh_nodes", "top_mbvh_nodes", "instance_indices", "global_ids",
                                          "tri_offsets")]


class RefBackend:
    """The reference's kernels behind the Backend method names (scene description replay) + ray casting / rendering."""

    def __init__(self):
        self.L = lib()
        self.o = orc.OracleBackend(det_eps=1e-4)   # builder + scene bookkeeping only
        self.OL = self.o.L
        self.OL.orc_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._keep = {}

    # ---- Backend surface: forwarded to the oracle's scene bookkeeping; lights / materials are bound as submitted ----
    def set_3d_mesh(self, *a, **k):
        self.o.set_3d_mesh(*a, **k)

    def unload_3d_meshes(self, ids):
        self.o.unload_3d_meshes(ids)

    def set_3d_instances(self, *a, **k):
        self.o.set_3d_instances(*a, **k)

    def set_skins(self, *a, **k):
        self.o.set_skins(*a, **k)

    def set_materials(self, m, changed=None):
        self._keep["materials"] = np.ascontiguousarray(m).copy()
        self.o.set_materials(m)

    def set_textures(self, textures=None, changed=None):
        self.o.set_textures(textures)

    def set_skybox(self, skybox=None):
        self.o.set_skybox(skybox)

    def _light(self, name, l):
        self._keep[name] = np.ascontiguousarray(l).copy()
        getattr(self.o, "set_" + name)(l)

    def set_area_lights(self, l, changed=None):
        self._light("area_lights", l)

    def set_point_lights(self, l, changed=None):
        self._light("point_lights", l)

    def set_spot_lights(self, l, changed=None):
        self._light("spot_lights", l)

    def set_directional_lights(self, l, changed=None):
        self._light("directional_lights", l)

    def synchronize(self):
        self.o.synchronize()
        counts = np.zeros(8, dtype=np.uint64)
        self.OL.orc_flatten(self.o.h, _ptr(counts), None)
        nt, npi, nb, nm, ni, ntb, ntm, nti = (int(c) for c in counts)
        f = {
            "triangles": np.zeros(max(1, nt) * 176, np.uint8), "prim_indices": np.zeros(max(1, npi), np.uint32), "bvh_nodes": np.zeros(max(1, nb) * 32, np.uint8),
            "mbvh_nodes": np.zeros(max(1, nm) * 128, np.uint8), "instances": np.zeros(max(1, ni) * 256, np.uint8), "top_bvh_nodes": np.zeros(max(1, ntb) * 32, np.uint8),
            "top_mbvh_nodes": np.zeros(max(1, ntm) * 128, np.uint8), "instance_indices": np.zeros(max(1, nti), np.uint32), "global_ids": np.zeros(max(1, ni), np.int32),
            "tri_offsets": np.zeros(max(1, ni), np.uint32),
        }
        flat = _Flat(*[f[n].ctypes.data for n, _ in _Flat._fields_])
        self.OL.orc_flatten(self.o.h, _ptr(counts), C.byref(flat))
        self._keep["flat"] = f
        self.n_instances = ni
        self.global_ids, self.tri_offsets = f["global_ids"][:ni], f["tri_offsets"][:ni]
        self.L.ref_bind_scene(_ptr(f["triangles"]), _ptr(f["prim_indices"]), _ptr(f["bvh_nodes"]), _ptr(f["mbvh_nodes"]), _ptr(f["instances"]), _ptr(f["instance_indices"]),
                              _ptr(f["top_bvh_nodes"]), _ptr(f["top_mbvh_nodes"]))
        k = self._keep
        empty = np.zeros(96, np.uint8)
        self.L.ref_bind_materials(_ptr(k.get("materials", empty)))
        lights = [k.get(n) for n in ("area_lights", "point_lights", "spot_lights", "directional_lights")]
        args = []
        for l in lights:
            args += [_ptr(l if l is not None and len(l) else empty), 0 if l is None else len(l)]
        self.L.ref_bind_lights(*args)
        # texture units: the oracle's software sampler (fixed-function fetches are not shader source)
        self.L.ref_set_texture_callback(C.cast(self.OL.orc_texture_callback, C.c_void_p), self.o.h)

    # ---- the reference's kernels -------------------------------------------------------------------------------
    def _map_hits(self, hits):
        out = hits.copy()
        hit = hits["inst"] >= 0
        if self.n_instances:
            idx = np.where(hit, hits["inst"], 0)
            out["inst"] = np.where(hit, self.global_ids[idx], -1)
            out["prim"] = np.where(hit, hits["prim"] - self.tri_offsets[idx].astype(np.int64), -1).astype(np.int32)
        return out

    def trace_closest(self, rays, mode=0):
        """ray_gen.comp:310-362 (mode 0, MBVH — what the shaders run) or :253-308 (mode 1, BVH2); hits mapped to (global instance id, mesh-local prim)."""
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), dtype=orc.HIT)
        if self.n_instances == 0:
            hits["inst"] = -1; hits["prim"] = -1; hits["t"] = rays["tmax"]
            return hits
        self.L.ref_trace_closest(_ptr(rays), len(rays), _ptr(hits), mode)
        return self._map_hits(hits)

    def trace_any(self, rays, mode=0):
        rays = np.ascontiguousarray(rays)
        occ = np.zeros(len(rays), dtype=np.uint32)
        if self.n_instances:
            self.L.ref_trace_any(_ptr(rays), len(rays), _ptr(occ), mode)
        return occ

    def render(self, view, w, h, spp, depth=3, clamp=10.0, first_sample=256, acc=None):
        """`spp` frames of RayTracer::render (lib.rs:1685-1729) starting at sample index first_sample (>= 256: hash RNG;
        below: the blue-noise tables, which must have been set).  Returns (acc[h,w,4], image[h,w,4] = blit.comp output, counters)."""
        if acc is None:
            acc = np.zeros((h, w, 4), dtype=np.float32)
        img = np.zeros((h, w, 4), dtype=np.float32)
        ctr = np.zeros(4, dtype=np.uint64)
        v = np.ascontiguousarray(view)
        self.L.ref_render(_ptr(v), w, h, first_sample, spp, depth, clamp, _ptr(acc), _ptr(img), _ptr(ctr))
        return acc, img, {"shaded": int(ctr[0]), "extension_rays": int(ctr[1]), "shadow_rays": int(ctr[2])}

    def set_blue_noise(self, table):
        t = np.ascontiguousarray(table, dtype=np.int32)
        self.L.ref_set_blue_noise(_ptr(t), len(t))
