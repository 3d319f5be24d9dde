"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY (pinned against the reference's own shader sources: see the header of oracle.cpp).  May be imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; nothing under
rfw_rs_b200/ imports it.  Exposes the same method names as the backend boundary
(crates/rfw-backend/src/lib.rs:35-82) so a scene description can be replayed on both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RFWB200_ORACLE_LIB: another build of the same source (bench.py's CPU arm compiles one with -march=native on the box it runs on)
_LIB_PATH = os.environ.get("RFWB200_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")

MODE_MBVH, MODE_BVH2, MODE_BRUTE = 0, 1, 2

RAY = np.dtype([("origin", np.float32, 3), ("tmin", np.float32), ("direction", np.float32, 3), ("tmax", np.float32)])
HIT = np.dtype([("inst", np.int32), ("prim", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])


def build(force=False):
    if os.environ.get("RFWB200_ORACLE_LIB"):
        return _LIB_PATH
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_mesh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_unload_mesh.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_instances.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        for nm in ("materials", "area_lights", "point_lights", "spot_lights", "directional_lights"):
            getattr(L, "orc_set_" + nm).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_set_mesh_skin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_set_instance_skins.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_set_num_skins.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_skin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_get_skinned_triangles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_get_skinned_triangles.restype = C.c_uint32
        L.orc_set_num_textures.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_set_skybox.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_sample_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_build.argtypes = [C.c_void_p]
        L.orc_build.restype = C.c_double
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orc_trace_closest.restype = C.c_double
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.orc_trace_any.restype = C.c_double
        L.orc_primary_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint32] * 9 + [C.c_float, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_debug_view.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
        L.orc_path_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_triangle_test.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_triangle_test.restype = C.c_int
        L.orc_wang_hash.argtypes = [C.c_uint32]
        L.orc_wang_hash.restype = C.c_uint32
        L.orc_randf.argtypes = [C.c_void_p]
        L.orc_randf.restype = C.c_float
        L.orc_random_barycentrics.argtypes = [C.c_float, C.c_void_p]
        L.orc_safe_origin.argtypes = [C.c_void_p] * 4
        L.orc_bvh_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_num_live_instances.argtypes = [C.c_void_p]
        L.orc_num_live_instances.restype = C.c_uint32
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleBackend:
    """CPU oracle behind the Backend method names (set_3d_mesh ... synchronize) plus the ray-casting
    extensions.  det_eps: Möller–Trumbore determinant epsilon (reference 1e-4 GLSL / 1e-6 Rust twin)."""

    def __init__(self, det_eps=0.0, mode=MODE_MBVH, threads=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.det_eps = float(det_eps)
        self.mode = mode
        self.threads = threads
        self.build_seconds = 0.0
        self.trace_seconds = 0.0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Backend surface -----------------------------------------------------------------------
    def set_3d_mesh(self, mesh_id, triangles, vertices=None, flags=3, skin_data=None):
        t = np.ascontiguousarray(triangles)
        assert t.dtype.itemsize == 176
        self.L.orc_set_mesh(self.h, mesh_id, _ptr(t), len(t))
        sk = np.zeros(0, dtype=np.uint8) if skin_data is None else np.ascontiguousarray(skin_data)
        assert len(sk) == 0 or sk.dtype.itemsize == 32
        self.L.orc_set_mesh_skin(self.h, mesh_id, _ptr(sk), len(sk))

    def unload_3d_meshes(self, ids):
        for i in ids:
            self.L.orc_unload_mesh(self.h, int(i))

    def set_3d_instances(self, mesh_id, matrices, skin_ids=None, flags=None):
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 16)
        self.L.orc_set_instances(self.h, mesh_id, _ptr(m), len(m))
        sk = np.zeros(0, dtype=np.int32) if skin_ids is None else np.ascontiguousarray(skin_ids, dtype=np.int32)
        self.L.orc_set_instance_skins(self.h, mesh_id, _ptr(sk), len(sk))

    def skinned_triangles(self, mesh_id, index, dtype):
        """RTTriangle records of the skinned copy of instance (mesh, index) (valid after synchronize)."""
        n = self.L.orc_get_skinned_triangles(self.h, mesh_id, index, None)
        out = np.zeros(n, dtype=dtype)
        if n:
            self.L.orc_get_skinned_triangles(self.h, mesh_id, index, _ptr(out))
        return out

    def set_blue_noise(self, table=None):
        t = np.zeros(0, dtype=np.uint32) if table is None else np.ascontiguousarray(table, dtype=np.uint32)
        self.L.orc_set_blue_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        self.L.orc_set_blue_noise(self.h, _ptr(t), len(t))

    def set_skins(self, skins=(), changed=None):
        self.L.orc_set_num_skins(self.h, len(skins))
        for k, j in enumerate(skins):
            a = np.ascontiguousarray(j, dtype=np.float32).reshape(-1, 16)
            self.L.orc_set_skin(self.h, k, _ptr(a), len(a))

    def _set(self, name, arr, size):
        a = np.ascontiguousarray(arr)
        assert a.dtype.itemsize == size or len(a) == 0
        getattr(self.L, "orc_set_" + name)(self.h, _ptr(a), len(a))

    def set_materials(self, m):
        self._set("materials", m, 96)

    def set_textures(self, textures=None, changed=None):
        textures = list(textures or [])
        self.L.orc_set_num_textures(self.h, len(textures))
        for i, t in enumerate(textures):
            b = np.ascontiguousarray(t.bytes, dtype=np.uint8)
            self.L.orc_set_texture(self.h, i, t.width, t.height, t.mip_levels, t.format, _ptr(b), b.size)

    def set_skybox(self, skybox=None):
        if skybox is None:
            self.L.orc_set_skybox(self.h, 0, 0, 0, 0, None, 0)
        else:
            b = np.ascontiguousarray(skybox.bytes, dtype=np.uint8)
            self.L.orc_set_skybox(self.h, skybox.width, skybox.height, skybox.mip_levels, skybox.format, _ptr(b), b.size)

    def sample_texture(self, tex, mode, u, v, lod):
        """Known-answer access to the samplers: tex index (-1 = skybox); mode 0 fetchTexel(level), 1 trilinear(lambda), 2 skybox level."""
        out = np.zeros(4, dtype=np.float32)
        self.L.orc_sample_texture(self.h, tex, mode, u, v, lod, _ptr(out))
        return out

    def set_area_lights(self, l):
        self._set("area_lights", l, 96)

    def set_point_lights(self, l):
        self._set("point_lights", l, 32)

    def set_spot_lights(self, l):
        self._set("spot_lights", l, 48)

    def set_directional_lights(self, l):
        self._set("directional_lights", l, 32)

    def synchronize(self):
        self.build_seconds = self.L.orc_build(self.h)

    # ---- ray casting -----------------------------------------------------------------------------
    def trace_closest(self, rays, mode=None, det_eps=None, counters=False):
        rays = np.ascontiguousarray(rays)
        hits = np.empty(len(rays), dtype=HIT)
        ctr = np.zeros(2, dtype=np.uint64)
        self.trace_seconds = self.L.orc_trace_closest(
            self.h, _ptr(rays), len(rays), _ptr(hits), self.det_eps if det_eps is None else det_eps,
            self.mode if mode is None else mode, self.threads, _ptr(ctr) if counters else None)
        return (hits, ctr) if counters else hits

    def trace_any(self, rays, mode=None, det_eps=None):
        rays = np.ascontiguousarray(rays)
        occ = np.empty(len(rays), dtype=np.uint32)
        self.trace_seconds = self.L.orc_trace_any(self.h, _ptr(rays), len(rays), _ptr(occ), self.det_eps if det_eps is None else det_eps,
                                                  self.mode if mode is None else mode, self.threads)
        return occ

    def primary_rays(self, view, w, h):
        rays = np.empty(w * h, dtype=RAY)
        v = np.ascontiguousarray(view)
        self.L.orc_primary_rays(_ptr(v), w, h, _ptr(rays))
        return rays

    def render(self, view, w, h, spp, depth, clamp=10.0, sky=(0, 0, 0), first_sample=0, window=None, acc=None):
        """Accumulate `spp` frames; returns (acc[h,w,4], stats dict)."""
        if acc is None:
            acc = np.zeros((h, w, 4), dtype=np.float32)
        x0, y0, x1, y1 = (0, 0, w, h) if window is None else window
        v = np.ascontiguousarray(view)
        skya = np.asarray(sky, dtype=np.float32)
        st = np.zeros(4, dtype=np.uint64)
        secs = self.L.orc_render(self.h, _ptr(v), w, h, x0, y0, x1, y1, first_sample, spp, depth, clamp, _ptr(skya), self.det_eps,
                                 self.threads, _ptr(acc), _ptr(st))
        return acc, {"samples": int(st[0]), "extension_rays": int(st[1]), "shadow_rays": int(st[2]), "segments": int(st[3]), "seconds": secs}

    def debug_view(self, view, w, h, mode):
        """RenderMode debug views at the primary hit: 1 normal, 2 albedo | material id, 3 world position | t."""
        out = np.zeros((h, w, 4), dtype=np.float32)
        v = np.ascontiguousarray(view)
        self.L.orc_debug_view(self.h, _ptr(v), w, h, mode, self.det_eps, _ptr(out))
        return out

    def path_probe(self, view, w, h, path_id, sample, depth, clamp=10.0, sky=(0, 0, 0)):
        """Debug: per-segment state of one path: rows of [O(3) D(3) inst prim t u v T(3) pdf 1 | nextO(3) nextD(3) acc.x newPdf]."""
        out = np.zeros((depth, 24), dtype=np.float32)
        v = np.ascontiguousarray(view)
        skya = np.asarray(sky, dtype=np.float32)
        self.L.orc_path_probe(self.h, _ptr(v), w, h, path_id, sample, depth, clamp, _ptr(skya), self.det_eps, _ptr(out))
        return out

    def max_threads(self):
        return self.L.orc_max_threads()
