"""Wire structs of the rfw-rs backend boundary as numpy dtypes / ctypes structures.

Byte-identical to include/rfwb200.h, which mirrors the reference's ``#[repr(C)]`` types
(crates/rfw-backend/src/structs.rs:879-918, 251-267, 369-394, 484-515; lights.rs:6-30, 100-108,
199-209, 293-301).  tests/test_layout.py checks sizes/offsets against the C header, the same
contract as the reference's ABI test (backends/metal/src/lib.rs:270-348).
"""
import ctypes as C

import numpy as np

f4 = np.float32

AABB = np.dtype([("min", f4, 3), ("extra1", np.int32), ("max", f4, 3), ("extra2", np.int32)])

RT_TRIANGLE = np.dtype(
    [
        ("vertex0", f4, 3), ("u0", f4),
        ("vertex1", f4, 3), ("u1", f4),
        ("vertex2", f4, 3), ("u2", f4),
        ("normal", f4, 3), ("v0", f4),
        ("n0", f4, 3), ("v1", f4),
        ("n1", f4, 3), ("v2", f4),
        ("n2", f4, 3), ("id", np.int32),
        ("tangent0", f4, 4),
        ("tangent1", f4, 4),
        ("tangent2", f4, 4),
        ("light_id", np.int32), ("mat_id", np.int32), ("lod", f4), ("area", f4),
    ]
)

VERTEX3D = np.dtype(
    [("vertex", f4, 4), ("normal", f4, 3), ("mat_id", np.uint32), ("uv", f4, 2), ("pad0", f4), ("pad1", f4), ("tangent", f4, 4)]
)

VERTEX_MESH = np.dtype([("bounds", AABB), ("first", np.uint32), ("last", np.uint32), ("mat_id", np.uint32), ("padding", np.uint32)])

JOINT_DATA = np.dtype([("joint", np.uint32, 4), ("weight", f4, 4)])

DEVICE_MATERIAL = np.dtype(
    [
        ("color", f4, 4), ("absorption", f4, 4), ("specular", f4, 4), ("parameters", np.uint32, 4),
        ("flags", np.uint32), ("diffuse_map", np.int32), ("normal_map", np.int32), ("metallic_roughness_map", np.int32),
        ("emissive_map", np.int32), ("sheen_map", np.int32), ("_dummy", np.int32, 2),
    ]
)

CAMERA_VIEW3D = np.dtype(
    [
        ("pos", f4, 3), ("right", f4, 3), ("up", f4, 3), ("p1", f4, 3), ("direction", f4, 3), ("lens_size", f4),
        ("spread_angle", f4), ("epsilon", f4), ("inv_width", f4), ("inv_height", f4),
        ("near_plane", f4), ("far_plane", f4), ("aspect_ratio", f4), ("fov", f4),
        ("custom0", f4, 4), ("custom1", f4, 4),
    ]
)

AREA_LIGHT = np.dtype(
    [
        ("position", f4, 3), ("energy", f4), ("normal", f4, 3), ("area", f4),
        ("vertex0", f4, 3), ("inst_idx", np.int32), ("vertex1", f4, 3), ("mesh_id", np.int32),
        ("radiance", f4, 3), ("_dummy1", np.int32), ("vertex2", f4, 3), ("_dummy2", np.int32),
    ]
)
SPOT_LIGHT = np.dtype([("position", f4, 3), ("cos_inner", f4), ("radiance", f4, 3), ("cos_outer", f4), ("direction", f4, 3), ("energy", f4)])
POINT_LIGHT = np.dtype([("position", f4, 3), ("energy", f4), ("radiance", f4, 3), ("_dummy", f4)])
DIRECTIONAL_LIGHT = np.dtype([("direction", f4, 3), ("energy", f4), ("radiance", f4, 3), ("_dummy", f4)])

RAY = np.dtype([("origin", f4, 3), ("tmin", f4), ("direction", f4, 3), ("tmax", f4)])
HIT = np.dtype([("inst", np.int32), ("prim", np.int32), ("t", f4), ("u", f4), ("v", f4)])
# the reference's 16-byte hit record (PathState.state, ray_extend.comp:267): bary = u16(65535 u) | u16(65535 v) << 16
HIT_PACKED = np.dtype([("inst", np.int32), ("prim", np.int32), ("t", f4), ("bary", np.uint32)])


def unpack_hits(packed):
    """RfwHitPacked -> RfwHit with the reference's unpacking (shade.comp:41-46: (bary & 65535) / 65535, (bary >> 16) / 65535)."""
    h = np.zeros(len(packed), HIT)
    h["inst"], h["prim"], h["t"] = packed["inst"], packed["prim"], packed["t"]
    h["u"] = (packed["bary"] & np.uint32(65535)).astype(f4) / f4(65535.0)
    h["v"] = (packed["bary"] >> np.uint32(16)).astype(f4) / f4(65535.0)
    return h


# rtbvh RayPacket4 as rfw fills it (crates/rfw-backend/src/structs.rs:656-667): ten SoA lanes of four floats
RAY_PACKET4 = np.dtype([(n, f4, 4) for n in ("origin_x", "origin_y", "origin_z", "direction_x", "direction_y", "direction_z", "t",
                                             "inv_direction_x", "inv_direction_y", "inv_direction_z")])


def rays_to_packets4(rays):
    """Groups rays (wire.RAY, count a multiple of 4) into RAY_PACKET4 records; packet.t = ray.tmax."""
    n = len(rays) // 4
    pk = np.zeros(n, RAY_PACKET4)
    o = rays["origin"][: 4 * n].reshape(n, 4, 3); d = rays["direction"][: 4 * n].reshape(n, 4, 3)
    for k, a in enumerate("xyz"):
        pk["origin_" + a] = o[:, :, k]; pk["direction_" + a] = d[:, :, k]
        with np.errstate(divide="ignore"):
            pk["inv_direction_" + a] = np.float32(1.0) / d[:, :, k]
    pk["t"] = rays["tmax"][: 4 * n].reshape(n, 4)
    return pk

EXPECTED_SIZES = {
    "RfwAabb": (AABB, 32), "RfwRTTriangle": (RT_TRIANGLE, 176), "RfwVertex3D": (VERTEX3D, 64), "RfwVertexMesh": (VERTEX_MESH, 48),
    "RfwJointData": (JOINT_DATA, 32), "RfwDeviceMaterial": (DEVICE_MATERIAL, 96), "RfwCameraView3D": (CAMERA_VIEW3D, 128),
    "RfwAreaLight": (AREA_LIGHT, 96), "RfwSpotLight": (SPOT_LIGHT, 48), "RfwPointLight": (POINT_LIGHT, 32),
    "RfwDirectionalLight": (DIRECTIONAL_LIGHT, 32), "RfwRay": (RAY, 32), "RfwHit": (HIT, 20), "RfwRayPacket4": (RAY_PACKET4, 160), "RfwHitPacked": (HIT_PACKED, 16),
}
for _name, (_dt, _sz) in EXPECTED_SIZES.items():
    assert _dt.itemsize == _sz, (_name, _dt.itemsize, _sz)


# ---- ctypes mirrors of the FFI repack structs (include/rfwb200.h) ---------------------------------
class CAabb(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("extra1", C.c_int32), ("max", C.c_float * 3), ("extra2", C.c_int32)]


class CMeshData3D(C.Structure):
    _fields_ = [
        ("triangles", C.c_void_p), ("num_triangles", C.c_uint32),
        ("vertices", C.c_void_p), ("num_vertices", C.c_uint32),
        ("ranges", C.c_void_p), ("num_ranges", C.c_uint32),
        ("skin_data", C.c_void_p), ("num_skin_data", C.c_uint32),
        ("flags", C.c_uint32), ("bounds", CAabb),
    ]


class CInstancesData3D(C.Structure):
    _fields_ = [("matrices", C.c_void_p), ("skin_ids", C.c_void_p), ("flags", C.c_void_p), ("num_instances", C.c_uint32), ("local_aabb", CAabb)]


class CTextureData(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("mip_levels", C.c_uint32), ("bytes", C.c_void_p), ("num_bytes", C.c_uint64), ("format", C.c_uint32)]


class CConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("max_depth", C.c_uint32), ("clamp_value", C.c_float),
        ("tile_size", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("sky", C.c_float * 3), ("reserved", C.c_uint32 * 8),
    ]


class CBuildStats(C.Structure):
    _fields_ = [
        ("num_meshes", C.c_uint32), ("num_instances", C.c_uint32), ("num_triangles", C.c_uint64), ("blas_nodes", C.c_uint64),
        ("tlas_nodes", C.c_uint64), ("bvh_bytes", C.c_uint64), ("blas_build_ms", C.c_float), ("tlas_build_ms", C.c_float),
        ("upload_ms", C.c_float), ("sah_cost", C.c_float), ("checksum", C.c_uint64), ("tlas_depth", C.c_uint32), ("blas_depth", C.c_uint32),
    ]


class CTraceStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("instances_entered", C.c_uint64),
        ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("stack_overflows", C.c_uint32), ("reserved", C.c_uint32),
    ]


class CRenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("segments", C.c_uint64), ("render_ms", C.c_float),
                ("stage_ms", C.c_float * 5), ("stage_timing", C.c_uint32), ("stack_overflows", C.c_uint32),
                ("gather_ms", C.c_float), ("frame_ms", C.c_float)]


def stats_to_dict(s):
    return {name: (list(getattr(s, name)) if isinstance(getattr(s, name), C.Array) else getattr(s, name)) for name, *_ in s._fields_}


class CSkinData(C.Structure):
    """RfwSkinData (include/rfwb200.h): FFI repack of SkinData<'a>, crates/rfw-backend/src/structs.rs:6-11."""
    _fields_ = [("inverse_bind_matrices", C.c_void_p), ("joint_matrices", C.c_void_p), ("num_joints", C.c_uint32)]
