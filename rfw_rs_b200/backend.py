"""ctypes binding of librfwb200.so — the host-side mirror of the reference's backend interface.

``B200Backend`` exposes the methods of ``rfw_backend::Backend`` (crates/rfw-backend/src/lib.rs:35-82)
under the same names and argument meaning (numpy arrays of the ``#[repr(C)]`` wire structs replace Rust
slices), plus the ray-casting extensions of include/rfwb200.h.  Everything goes through the C ABI — this
file contains no computation.  The library is CUDA-only: constructing a backend without a B200 raises
``RfwError``; there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import wire

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RFWB200_LIB") or os.path.join(_HERE, "librfwb200.so")  # RFWB200_LIB: A/B builds in tuning scripts


class RfwError(RuntimeError):
    pass


def build_library(force=False):
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8"]
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "clean"])
    subprocess.check_call(args)
    return LIB_PATH


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RfwError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, f32p = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p
    sig = {
        "rfwb200_create": ([vp, C.POINTER(vp)], i32),
        "rfwb200_destroy": ([vp], None),
        "rfwb200_set_3d_mesh": ([vp, u32, vp], i32),
        "rfwb200_unload_3d_meshes": ([vp, vp, u32], i32),
        "rfwb200_set_3d_instances": ([vp, u32, vp], i32),
        "rfwb200_set_materials": ([vp, vp, u32, vp], i32),
        "rfwb200_set_textures": ([vp, vp, u32, vp], i32),
        "rfwb200_synchronize": ([vp], i32),
        "rfwb200_render": ([vp, vp, u32], i32),
        "rfwb200_resize": ([vp, u32, u32, C.c_double], i32),
        "rfwb200_set_point_lights": ([vp, vp, u32, vp], i32),
        "rfwb200_set_spot_lights": ([vp, vp, u32, vp], i32),
        "rfwb200_set_area_lights": ([vp, vp, u32, vp], i32),
        "rfwb200_set_directional_lights": ([vp, vp, u32, vp], i32),
        "rfwb200_set_skybox": ([vp, vp], i32),
        "rfwb200_set_skins": ([vp, vp, u32, vp], i32),
        "rfwb200_set_blue_noise": ([vp, vp, u32], i32),
        "rfwb200_set_2d_mesh": ([vp, u32, vp, u32, C.c_int32], i32),
        "rfwb200_set_2d_instances": ([vp, u32, vp, u32], i32),
        "rfwb200_trace_closest": ([vp, vp, u64, vp], i32),
        "rfwb200_trace_any": ([vp, vp, u64, vp], i32),
        "rfwb200_trace_closest_packed": ([vp, vp, u64, vp], i32),
        "rfwb200_trace_closest_packed_device": ([vp, vp, u64, vp, i32], i32),
        "rfwb200_intersect_t": ([vp, vp, u64, vp], i32),
        "rfwb200_depth_test": ([vp, vp, u64, vp, vp], i32),
        "rfwb200_intersect4": ([vp, vp, u64, vp, vp, vp], i32),
        "rfwb200_occludes4": ([vp, vp, u64, vp, vp], i32),
        "rfwb200_trace_closest_device": ([vp, vp, u64, vp, i32], i32),
        "rfwb200_trace_any_device": ([vp, vp, u64, vp, i32], i32),
        "rfwb200_trace_closest_counted": ([vp, vp, u64, vp, vp], i32),
        "rfwb200_cast_primary": ([vp, vp, vp], i32),
        "rfwb200_render_spp": ([vp, vp, u32, u32], i32),
        "rfwb200_reset_accumulator": ([vp], i32),
        "rfwb200_read_accumulator": ([vp, f32p], i32),
        "rfwb200_read_output": ([vp, f32p], i32),
        "rfwb200_export_tiles_device": ([vp, vp, u32, vp], i32),
        "rfwb200_assemble_tiles_device": ([vp, vp, u32, u32, vp], i32),
        "rfwb200_comm_unique_id": ([vp], i32),
        "rfwb200_comm_init": ([vp, vp, u32, u32], i32),
        "rfwb200_comm_destroy": ([vp], i32),
        "rfwb200_gather_image": ([vp, u32, vp], i32),
        "rfwb200_render_gather": ([vp, vp, u32, u32, u32, vp], i32),
        "rfwb200_nccl_version": ([], i32),
        "rfwb200_sample_count": ([vp], u32),
        "rfwb200_tiles_per_rank": ([vp], u32),
        "rfwb200_build_stats": ([vp, vp], i32),
        "rfwb200_trace_stats": ([vp, vp], i32),
        "rfwb200_render_stats": ([vp, vp], i32),
        "rfwb200_set_option": ([vp, C.c_char_p, C.c_int64], i32),
        "rfwb200_debug_read_queue": ([vp, u32, vp, vp, vp, vp, u32, vp], i32),
        "rfwb200_tile_layout": ([u32, u32, u32, vp, u32], u32),
        "rfwb200_measure_l2_read_gbs": ([vp, u64, u32, vp], i32),
        "rfwb200_host_alloc": ([u64], vp),
        "rfwb200_host_free": ([vp], None),
        "rfwb200_last_error": ([], C.c_char_p),
        "rfwb200_version": ([], C.c_char_p),
        "rfwb200_launch_count": ([vp], u64),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def _ptr(a):
    return None if a is None or len(a) == 0 else a.ctypes.data


class PinnedArray:
    """numpy view over pinned host memory from rfwb200_host_alloc (for the host-buffer entry points)."""

    def __init__(self, count, dtype):
        self.L = load_library()
        self.dtype = np.dtype(dtype)
        self.nbytes = max(1, count * self.dtype.itemsize)
        self.ptr = self.L.rfwb200_host_alloc(self.nbytes)
        if not self.ptr:
            raise RfwError("pinned allocation failed: " + self.L.rfwb200_last_error().decode())
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.rfwb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class B200Backend:
    """The B200 ray-tracing backend behind the Backend trait's method names."""

    def __init__(self, width=0, height=0, device=0, max_depth=3, clamp_value=10.0, tile_size=64, rank=0, world=1, sky=(0.0, 0.0, 0.0)):
        self.L = load_library()
        cfg = wire.CConfig()
        cfg.device, cfg.width, cfg.height, cfg.max_depth, cfg.clamp_value = device, width, height, max_depth, clamp_value
        cfg.tile_size, cfg.rank, cfg.world = tile_size, rank, world
        cfg.sky[0], cfg.sky[1], cfg.sky[2] = sky
        self.width, self.height = width, height
        self.h = C.c_void_p()
        rc = self.L.rfwb200_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RfwError(f"rfwb200_create failed ({rc}): {self.L.rfwb200_last_error().decode()}")

    # FromWindowHandle::init analogue (crates/rfw-backend/src/lib.rs:26-33); the window is ignored
    @classmethod
    def init(cls, window, width, height, scale=1.0, **kw):
        return cls(width, height, **kw)

    def close(self):
        if getattr(self, "h", None):
            self.L.rfwb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RfwError(f"{what} failed ({rc}): {self.L.rfwb200_last_error().decode()}")

    # ---- Backend trait ---------------------------------------------------------------------------
    def set_3d_mesh(self, mesh_id, triangles, vertices=None, flags=3, skin_data=None):
        t = np.ascontiguousarray(triangles)
        if len(t) and t.dtype.itemsize != 176:
            raise RfwError("set_3d_mesh: triangles must be 176-byte RTTriangle records")
        d = wire.CMeshData3D()
        d.triangles, d.num_triangles = _ptr(t), len(t)
        if vertices is not None and len(vertices):
            v = np.ascontiguousarray(vertices)
            d.vertices, d.num_vertices = _ptr(v), len(v)
        if skin_data is not None and len(skin_data):
            sk = np.ascontiguousarray(skin_data)
            if sk.dtype.itemsize != 32:
                raise RfwError("set_3d_mesh: skin_data must be 32-byte JointData records")
            d.skin_data, d.num_skin_data = _ptr(sk), len(sk)
        d.flags = flags
        self._ck(self.L.rfwb200_set_3d_mesh(self.h, mesh_id, C.addressof(d)), "set_3d_mesh")

    def unload_3d_meshes(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        self._ck(self.L.rfwb200_unload_3d_meshes(self.h, _ptr(a), len(a)), "unload_3d_meshes")

    def set_3d_instances(self, mesh_id, matrices, skin_ids=None, flags=None):
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 16)
        d = wire.CInstancesData3D()
        d.matrices, d.num_instances = _ptr(m), len(m)
        if skin_ids is not None:
            sk = np.ascontiguousarray(skin_ids, dtype=np.int32)
            if len(sk) != len(m):
                raise RfwError("set_3d_instances: one skin id per instance")
            d.skin_ids = _ptr(sk)
        self._ck(self.L.rfwb200_set_3d_instances(self.h, mesh_id, C.addressof(d)), "set_3d_instances")

    def _set_array(self, fn, arr, size, what):
        a = np.ascontiguousarray(arr)
        if len(a) and a.dtype.itemsize != size:
            raise RfwError(f"{what}: records must be {size} bytes")
        self._ck(getattr(self.L, fn)(self.h, _ptr(a), len(a), None), what)

    def set_materials(self, materials, changed=None):
        self._set_array("rfwb200_set_materials", materials, 96, "set_materials")

    @staticmethod
    def _texture_data(t):
        d = wire.CTextureData()
        b = np.ascontiguousarray(t.bytes, dtype=np.uint8)
        d.width, d.height, d.mip_levels, d.format = t.width, t.height, t.mip_levels, t.format
        d.bytes, d.num_bytes = _ptr(b), b.size
        return d, b

    def set_textures(self, textures=None, changed=None):
        """textures: objects with width/height/mip_levels/format/bytes (scenes.Texture)."""
        textures = list(textures or [])
        if not textures:
            self._ck(self.L.rfwb200_set_textures(self.h, None, 0, None), "set_textures")
            return
        packed = [self._texture_data(t) for t in textures]  # keeps the byte arrays alive for the call
        arr = (wire.CTextureData * len(packed))(*[p[0] for p in packed])
        ch = None if changed is None else np.ascontiguousarray(changed, dtype=np.uint32)
        self._ck(self.L.rfwb200_set_textures(self.h, C.addressof(arr), len(packed), None if ch is None else _ptr(ch)), "set_textures")

    def set_point_lights(self, lights, changed=None):
        self._set_array("rfwb200_set_point_lights", lights, 32, "set_point_lights")

    def set_spot_lights(self, lights, changed=None):
        self._set_array("rfwb200_set_spot_lights", lights, 48, "set_spot_lights")

    def set_area_lights(self, lights, changed=None):
        self._set_array("rfwb200_set_area_lights", lights, 96, "set_area_lights")

    def set_directional_lights(self, lights, changed=None):
        self._set_array("rfwb200_set_directional_lights", lights, 32, "set_directional_lights")

    def set_skybox(self, skybox=None):
        if skybox is None:
            self._ck(self.L.rfwb200_set_skybox(self.h, None), "set_skybox")
            return
        d, _keep = self._texture_data(skybox)
        self._ck(self.L.rfwb200_set_skybox(self.h, C.addressof(d)), "set_skybox")

    def set_blue_noise(self, table=None):
        """u32 sampler tables in create_blue_noise_buffer's layout (None: hash RNG for every sample)."""
        if table is None:
            self._ck(self.L.rfwb200_set_blue_noise(self.h, None, 0), "set_blue_noise")
            return
        t = np.ascontiguousarray(table, dtype=np.uint32)
        self._ck(self.L.rfwb200_set_blue_noise(self.h, _ptr(t), len(t)), "set_blue_noise")

    def set_skins(self, skins=(), changed=None):
        """skins: list of (n_joints, 16) column-major joint matrices (SkinData::joint_matrices)."""
        skins = [np.ascontiguousarray(j, dtype=np.float32).reshape(-1, 16) for j in skins]
        if not skins:
            self._ck(self.L.rfwb200_set_skins(self.h, None, 0, None), "set_skins")
            return
        arr = (wire.CSkinData * len(skins))()
        for k, j in enumerate(skins):
            arr[k].inverse_bind_matrices = None
            arr[k].joint_matrices = _ptr(j)
            arr[k].num_joints = len(j)
        ch = None if changed is None else np.ascontiguousarray(changed, dtype=np.uint32)
        self._ck(self.L.rfwb200_set_skins(self.h, C.addressof(arr), len(skins), None if ch is None else _ptr(ch)), "set_skins")

    def set_2d_mesh(self, mesh_id, data=None):
        self._ck(self.L.rfwb200_set_2d_mesh(self.h, mesh_id, None, 0, -1), "set_2d_mesh")

    def set_2d_instances(self, mesh_id, matrices=None):
        self._ck(self.L.rfwb200_set_2d_instances(self.h, mesh_id, None, 0), "set_2d_instances")

    def synchronize(self):
        self._ck(self.L.rfwb200_synchronize(self.h), "synchronize")

    def render(self, view_2d, view_3d, mode=0):
        v = np.ascontiguousarray(view_3d)
        self._ck(self.L.rfwb200_render(self.h, _ptr(v), int(mode)), "render")

    def resize(self, window_size, scale_factor=1.0):
        self.width, self.height = window_size
        self._ck(self.L.rfwb200_resize(self.h, window_size[0], window_size[1], scale_factor), "resize")

    # ---- ray-casting / measurement extensions -------------------------------------------------------
    def trace_closest(self, rays, out=None):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 32
        hits = np.empty(len(rays), dtype=wire.HIT) if out is None else out
        self._ck(self.L.rfwb200_trace_closest(self.h, _ptr(rays), len(rays), _ptr(hits)), "trace_closest")
        return hits

    def trace_closest_packed(self, rays, out=None):
        """Closest hits as 16-byte wire.HIT_PACKED records (the reference's own hit record; wire.unpack_hits expands them)."""
        rays = np.ascontiguousarray(rays)
        hits = np.empty(len(rays), dtype=wire.HIT_PACKED) if out is None else out
        self._ck(self.L.rfwb200_trace_closest_packed(self.h, _ptr(rays), len(rays), _ptr(hits)), "trace_closest_packed")
        return hits

    def trace_closest_packed_device(self, d_rays_ptr, n, d_hits_ptr, sync=True):
        self._ck(self.L.rfwb200_trace_closest_packed_device(self.h, d_rays_ptr, n, d_hits_ptr, int(sync)), "trace_closest_packed_device")

    def trace_any(self, rays, out=None):
        rays = np.ascontiguousarray(rays)
        occ = np.empty(len(rays), dtype=np.uint32) if out is None else out
        self._ck(self.L.rfwb200_trace_any(self.h, _ptr(rays), len(rays), _ptr(occ)), "trace_any")
        return occ

    # ---- the rest of TIntersector (crates/rfw-scene/src/intersector.rs:77-166) --------------------------------------
    def intersect_t(self, rays):
        """Closest-hit distance per ray; -1 where the reference returns None."""
        rays = np.ascontiguousarray(rays)
        t = np.empty(len(rays), np.float32)
        self._ck(self.L.rfwb200_intersect_t(self.h, _ptr(rays), len(rays), _ptr(t)), "intersect_t")
        return t

    def depth_test(self, rays):
        """(t, depth): closest t (ray.tmax on a miss) and the number of acceleration-structure nodes the ray visited."""
        rays = np.ascontiguousarray(rays)
        t = np.empty(len(rays), np.float32)
        depth = np.empty(len(rays), np.uint32)
        self._ck(self.L.rfwb200_depth_test(self.h, _ptr(rays), len(rays), _ptr(t), _ptr(depth)), "depth_test")
        return t, depth

    def intersect4(self, packets, t_min=(1e-4,) * 4):
        """packets: array of wire.RAY_PACKET4 (updated in place: t of the lanes that hit).  Returns (inst[n, 4], prim[n, 4])."""
        assert packets.dtype.itemsize == 160 and packets.flags["C_CONTIGUOUS"]
        tm = np.asarray(t_min, np.float32)
        inst = np.empty((len(packets), 4), np.int32)
        prim = np.empty((len(packets), 4), np.int32)
        self._ck(self.L.rfwb200_intersect4(self.h, _ptr(packets), len(packets), _ptr(tm), _ptr(inst), _ptr(prim)), "intersect4")
        return inst, prim

    def occludes4(self, packets, t_min=(1e-3,) * 4):
        assert packets.dtype.itemsize == 160 and packets.flags["C_CONTIGUOUS"]
        tm = np.asarray(t_min, np.float32)
        occ = np.empty((len(packets), 4), np.uint32)
        self._ck(self.L.rfwb200_occludes4(self.h, _ptr(packets), len(packets), _ptr(tm), _ptr(occ)), "occludes4")
        return occ

    def trace_closest_device(self, d_rays_ptr, n, d_hits_ptr, sync=True):
        self._ck(self.L.rfwb200_trace_closest_device(self.h, d_rays_ptr, n, d_hits_ptr, int(sync)), "trace_closest_device")

    def trace_any_device(self, d_rays_ptr, n, d_occ_ptr, sync=True):
        self._ck(self.L.rfwb200_trace_any_device(self.h, d_rays_ptr, n, d_occ_ptr, int(sync)), "trace_any_device")

    def trace_closest_counted(self, d_rays_ptr, n, d_hits_ptr):
        st = wire.CTraceStats()
        self._ck(self.L.rfwb200_trace_closest_counted(self.h, d_rays_ptr, n, d_hits_ptr, C.addressof(st)), "trace_closest_counted")
        return wire.stats_to_dict(st)

    def cast_primary(self, view):
        v = np.ascontiguousarray(view)
        hits = np.empty(self.width * self.height, dtype=wire.HIT)
        self._ck(self.L.rfwb200_cast_primary(self.h, _ptr(v), _ptr(hits)), "cast_primary")
        return hits

    def render_spp(self, view, spp, depth=0):
        v = np.ascontiguousarray(view)
        self._ck(self.L.rfwb200_render_spp(self.h, _ptr(v), spp, depth), "render_spp")

    def reset_accumulator(self):
        self._ck(self.L.rfwb200_reset_accumulator(self.h), "reset_accumulator")

    def read_accumulator(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.L.rfwb200_read_accumulator(self.h, _ptr(out)), "read_accumulator")
        return out

    def read_output(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.L.rfwb200_read_output(self.h, _ptr(out)), "read_output")
        return out

    def export_tiles_device(self, d_ptr, capacity_tiles):
        n = C.c_uint32(0)
        self._ck(self.L.rfwb200_export_tiles_device(self.h, d_ptr, capacity_tiles, C.addressof(n)), "export_tiles_device")
        return n.value

    def assemble_tiles_device(self, d_gathered_ptr, tiles_per_rank, world, d_image_ptr):
        self._ck(self.L.rfwb200_assemble_tiles_device(self.h, d_gathered_ptr, tiles_per_rank, world, d_image_ptr), "assemble_tiles_device")

    # ---- multi-GPU: NCCL accumulator gather inside the library (rfwb200.h, "multi-GPU") ---------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId (rank 0 calls this and distributes them)."""
        L = load_library()
        buf = (C.c_uint8 * 128)()
        rc = L.rfwb200_comm_unique_id(C.addressof(buf))
        if rc != 0:
            raise RfwError(f"comm_unique_id failed ({rc}): {L.rfwb200_last_error().decode()}")
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.rfwb200_comm_init(self.h, C.addressof(buf), rank, world), "comm_init")

    def comm_destroy(self):
        self._ck(self.L.rfwb200_comm_destroy(self.h), "comm_destroy")

    def gather_image(self, root=0, d_image_ptr=None):
        """Collective: every rank calls it after render_spp.  root < world: that rank receives; root >= world: all do."""
        self._ck(self.L.rfwb200_gather_image(self.h, root, d_image_ptr), "gather_image")

    def render_gather(self, view, spp, depth=0, root=0, d_image_ptr=None):
        v = np.ascontiguousarray(view)
        self._ck(self.L.rfwb200_render_gather(self.h, _ptr(v), spp, depth, root, d_image_ptr), "render_gather")

    @property
    def sample_count(self):
        return self.L.rfwb200_sample_count(self.h)

    @property
    def tiles_per_rank(self):
        return self.L.rfwb200_tiles_per_rank(self.h)

    def build_stats(self):
        s = wire.CBuildStats()
        self._ck(self.L.rfwb200_build_stats(self.h, C.addressof(s)), "build_stats")
        return wire.stats_to_dict(s)

    def trace_stats(self):
        s = wire.CTraceStats()
        self._ck(self.L.rfwb200_trace_stats(self.h, C.addressof(s)), "trace_stats")
        return wire.stats_to_dict(s)

    def render_stats(self):
        s = wire.CRenderStats()
        self._ck(self.L.rfwb200_render_stats(self.h, C.addressof(s)), "render_stats")
        return wire.stats_to_dict(s)

    def set_option(self, key, value):
        self._ck(self.L.rfwb200_set_option(self.h, key.encode(), int(value)), "set_option")

    def debug_read_queue(self, which, capacity):
        """Debug: (O, D, T, S, count) of wavefront queue `which` after the last render call."""
        O = np.zeros((capacity, 4), np.float32); D = np.zeros((capacity, 4), np.float32)
        T = np.zeros((capacity, 4), np.float32); S = np.zeros((capacity, 4), np.float32)
        n = C.c_uint32(0)
        self._ck(self.L.rfwb200_debug_read_queue(self.h, which, _ptr(O), _ptr(D), _ptr(T), _ptr(S), capacity, C.addressof(n)), "debug_read_queue")
        return O, D, T, S, n.value

    def measure_l2_read_gbs(self, nbytes=32 << 20, iters=50):
        out = C.c_float(0)
        self._ck(self.L.rfwb200_measure_l2_read_gbs(self.h, nbytes, iters, C.addressof(out)), "measure_l2_read_gbs")
        return out.value

    def save_ppm(self, path):
        """Presentation stand-in for the reference's swap-chain blit (backends/gpu-rt/src/lib.rs:1753-1776): the output
        buffer (sqrt(acc / spp), blit.comp:22) as an 8-bit binary PPM."""
        img = np.clip(self.read_output()[..., :3], 0.0, 1.0)
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (self.width, self.height))
            f.write((img * 255.0 + 0.5).astype(np.uint8).tobytes())

    def launch_count(self):
        return self.L.rfwb200_launch_count(self.h)
