"""Synthetic scene / ray generators for the configs of BASELINE.json.

Host-side restatements of the rfw-scene producers that feed the backend boundary, so the backend can
be driven without Rust:
  * RTTriangle fill rules       crates/rfw-scene/src/objects_3d/mod.rs:331-383, structs.rs:970-983
  * Camera3D::get_view          crates/rfw-scene/src/camera/mod.rs:77-115, 246-252
  * into_device_material        crates/rfw-scene/src/material/list.rs:755-814
  * AreaLight::new              crates/rfw-backend/src/lights.rs:71-97
  * Quad3D / Sphere (icosphere) crates/rfw-scene/src/objects_3d/quad.rs:51-74, sphere.rs:11-24
All randomness is counter-based splitmix64 (identical in Python / C++ / CUDA): scenes and rays are
regenerated on the GPU box, never shipped.
"""
import numpy as np

from . import wire

MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_SCENE, SEED_RAYS, SEED_LIGHTS = 1234, 5678, 91011


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(seed, index):
    """u01 = (splitmix64(seed*0x9E3779B97F4A7C15 + index) >> 40) * 2^-24."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        x = splitmix64(base + np.asarray(index, dtype=np.uint64))
    return ((x >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)).astype(np.float32)


def _norm(v):
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-30)


def make_triangles(v0, v1, v2, mat_id=0, n0=None, n1=None, n2=None, t0=None, t1=None, t2=None):
    """Fill RTTriangle records the way Mesh3D::new does (objects_3d/mod.rs:331-383): unit geometric
    normal, vertex normals (flat unless given), id = index, Heron area, light_id = -1."""
    v0 = np.asarray(v0, np.float32); v1 = np.asarray(v1, np.float32); v2 = np.asarray(v2, np.float32)
    n = v0.shape[0]
    cr = np.cross((v1 - v0).astype(np.float64), (v2 - v0).astype(np.float64))
    ln = np.linalg.norm(cr, axis=1)
    keep = ln > 0
    if not keep.all():  # degenerate triangles give a NaN normal: filter
        v0, v1, v2, cr, ln = v0[keep], v1[keep], v2[keep], cr[keep], ln[keep]
        if n0 is not None:
            n0, n1, n2 = n0[keep], n1[keep], n2[keep]
        if t0 is not None:
            t0, t1, t2 = t0[keep], t1[keep], t2[keep]
        if np.ndim(mat_id) > 0:
            mat_id = np.asarray(mat_id)[keep]
        n = v0.shape[0]
    normal = (cr / ln[:, None]).astype(np.float32)
    t = np.zeros(n, dtype=wire.RT_TRIANGLE)
    t["vertex0"], t["vertex1"], t["vertex2"] = v0, v1, v2
    t["normal"] = normal
    t["n0"] = normal if n0 is None else n0
    t["n1"] = normal if n1 is None else n1
    t["n2"] = normal if n2 is None else n2
    t["id"] = np.arange(n, dtype=np.int32)
    # tangent: unit vector perpendicular to the normal, w = 1 (objects_3d/mod.rs:256-266)
    ref = np.where(np.abs(normal[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    tan = _norm(np.cross(normal, ref)).astype(np.float32)
    for k, tv in (("tangent0", t0), ("tangent1", t1), ("tangent2", t2)):
        t[k][:, :3] = tan if tv is None else tv  # per-vertex tangents keep shading continuous across shared edges
        t[k][:, 3] = 1.0
    t["light_id"] = -1
    t["mat_id"] = mat_id
    a = np.linalg.norm((v1 - v0).astype(np.float64), axis=1)
    b = np.linalg.norm((v2 - v1).astype(np.float64), axis=1)
    c = np.linalg.norm((v0 - v2).astype(np.float64), axis=1)
    s = (a + b + c) * 0.5
    t["area"] = np.sqrt(np.maximum(s * (s - a) * (s - b) * (s - c), 0.0)).astype(np.float32)  # Heron, structs.rs:977-983
    return t


def soup(n, s, seed=SEED_SCENE, mat_id=0, chunk=1 << 20):
    """Random triangle soup: centre ~ U[0,1)^3, vertices centre + U(-s,s)^3."""
    out = []
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = (np.arange(lo, hi, dtype=np.uint64)[:, None] * np.uint64(12) + np.arange(12, dtype=np.uint64)[None, :])
        r = u01(seed, idx)
        c = r[:, 0:3]
        d = (r[:, 3:12] * 2.0 - 1.0) * np.float32(s)
        v0 = c + d[:, 0:3]; v1 = c + d[:, 3:6]; v2 = c + d[:, 6:9]
        out.append(make_triangles(v0, v1, v2, mat_id))
    t = np.concatenate(out)
    t["id"] = np.arange(len(t), dtype=np.int32)
    return t


def random_rays(n, seed=SEED_RAYS, tmin=1e-4, tmax=1e26, lo=0.0, hi=1.0, start=0, chunk=1 << 22):
    """Incoherent rays: origin ~ U[lo,hi)^3, direction uniform on the sphere (z = 1-2u1, phi = 2 pi u2)."""
    rays = np.empty(n, dtype=wire.RAY)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        idx = (np.arange(start + a, start + b, dtype=np.uint64)[:, None] * np.uint64(5) + np.arange(5, dtype=np.uint64)[None, :])
        r = u01(seed, idx).astype(np.float64)
        z = 1.0 - 2.0 * r[:, 3]
        phi = 2.0 * np.pi * r[:, 4]
        rad = np.sqrt(np.maximum(0.0, 1.0 - z * z))
        rays["origin"][a:b] = (lo + (hi - lo) * r[:, 0:3]).astype(np.float32)
        rays["direction"][a:b] = np.stack([rad * np.cos(phi), rad * np.sin(phi), z], axis=1).astype(np.float32)
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


def icosphere(subdiv=3, radius=0.4, mat_id=0):
    """Icosphere, 20*4^subdiv triangles, smooth normals (Sphere Quality::High = 3, sphere.rs:11-24)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [tuple(np.array(v, np.float64) / np.linalg.norm(v)) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
             (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.array(verts[a]) + np.array(verts[b])) * 0.5
                verts.append(tuple(m / np.linalg.norm(m)))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    v = np.array(verts, np.float64)
    f = np.array(faces, np.int64)
    nrm = v.astype(np.float32)
    p = (v * radius).astype(np.float32)
    # per-vertex tangent = normalize(up x n) (Mesh3D::new derives per-vertex tangents too, objects_3d/mod.rs:256-266)
    up = np.where(np.abs(v[:, 1:2]) > 0.99, np.array([[1.0, 0.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))
    tanv = _norm(np.cross(up, v)).astype(np.float32)
    return make_triangles(p[f[:, 0]], p[f[:, 1]], p[f[:, 2]], mat_id, nrm[f[:, 0]], nrm[f[:, 1]], nrm[f[:, 2]], tanv[f[:, 0]], tanv[f[:, 1]], tanv[f[:, 2]])


def quad(pos, normal, width, height, mat_id=0):
    """Quad3D::generate_render_data (objects_3d/quad.rs:51-74): 2 triangles."""
    pos = np.asarray(pos, np.float64)
    n = np.asarray(normal, np.float64); n = n / np.linalg.norm(n)
    tmp = np.array([0.0, 1.0, 0.0]) if n[0] > 0.9 else np.array([1.0, 0.0, 0.0])
    tangent = 0.5 * width * _norm(np.cross(n, tmp))
    bitangent = 0.5 * height * np.cross(_norm(tangent), n)
    vs = np.array([pos - bitangent - tangent, pos + bitangent - tangent, pos - bitangent + tangent,
                   pos + bitangent - tangent, pos + bitangent + tangent, pos - bitangent + tangent], np.float32)
    t = make_triangles(vs[[0, 3]], vs[[1, 4]], vs[[2, 5]], mat_id)
    # the generator's winding does not always agree with `normal`; RTTriangle.normal follows the winding
    return t


def concat_meshes(parts):
    t = np.concatenate(parts)
    t["id"] = np.arange(len(t), dtype=np.int32)
    return t


def identity():
    return np.eye(4, dtype=np.float32)


def trs(translate=(0, 0, 0), rot_axis=(0, 1, 0), rot_angle=0.0, scale=1.0):
    """Column-major Mat4 T*R*S flattened the way glam stores it (16 floats, column after column)."""
    ax = np.asarray(rot_axis, np.float64); ax = ax / np.linalg.norm(ax)
    c, s = np.cos(rot_angle), np.sin(rot_angle)
    x, y, z = ax
    R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
    sc = np.asarray(scale, np.float64) * np.ones(3)
    M = np.eye(4)
    M[:3, :3] = R * sc[None, :]
    M[:3, 3] = translate
    return M.astype(np.float32)


def to_column_major(mats):
    """(n,4,4) row-indexed matrices -> (n,16) column-major floats as the boundary expects."""
    mats = np.asarray(mats, np.float32).reshape(-1, 4, 4)
    return np.ascontiguousarray(mats.transpose(0, 2, 1).reshape(-1, 16))


def camera_view(pos, direction, width, height, fov_deg=40.0, aperture=1e-4, focal_distance=1.0, aspect=None):
    """Camera3D::get_view (camera/mod.rs:77-115) with calculate_matrix (:246-252)."""
    pos = np.asarray(pos, np.float32)
    f32 = np.float32
    z = _norm(np.asarray(direction, np.float32)).astype(f32)
    y = np.array([0, 1, 0], f32)
    x = _norm(np.cross(z, y)).astype(f32)
    y = _norm(np.cross(x, z)).astype(f32)
    aspect = f32(width / height if aspect is None else aspect)
    fov = f32(fov_deg)
    spread_angle = (fov * f32(np.pi) / f32(180.0)) * (f32(1.0) / f32(height))
    screen_size = f32(np.tan(fov * f32(0.5) / (f32(180.0) / f32(np.pi))))
    fd = f32(focal_distance)
    center = pos + fd * z
    p1 = center - screen_size * x * fd * aspect + screen_size * fd * y
    p2 = center + screen_size * x * fd * aspect + screen_size * fd * y
    p3 = center - screen_size * x * fd * aspect - screen_size * fd * y
    v = np.zeros(1, dtype=wire.CAMERA_VIEW3D)
    v["pos"] = pos; v["right"] = p2 - p1; v["up"] = p3 - p1; v["p1"] = p1; v["direction"] = z
    v["lens_size"] = aperture; v["spread_angle"] = spread_angle; v["epsilon"] = 1e-4
    v["inv_width"] = f32(1.0) / f32(width); v["inv_height"] = f32(1.0) / f32(height)
    v["near_plane"] = 1e-2; v["far_plane"] = 1e5; v["aspect_ratio"] = aspect; v["fov"] = np.radians(fov)
    return v


def material(color=(0.8, 0.8, 0.8), metallic=0.0, roughness=1.0, specular_f=0.5, subsurface=0.0, specular=(1, 1, 1), transmission=0.0, eta=1.0,
             clearcoat=0.0, clearcoat_gloss=0.0, diffuse_map=-1, normal_map=-1, specular_tint=0.0, anisotropic=0.0, sheen=0.0, sheen_tint=0.0,
             absorption=(0, 0, 0)):
    """into_device_material (material/list.rs:755-814): u8-packed Disney parameters, no textures."""
    def ch(f):
        return int(min(f * 255.0, 255.0)) & 255

    def pk(a, b, c, d):
        return ch(a) | (ch(b) << 8) | (ch(c) << 16) | (ch(d) << 24)

    m = np.zeros(1, dtype=wire.DEVICE_MATERIAL)
    m["color"][0, :3] = color; m["color"][0, 3] = 1.0
    m["specular"][0, :3] = specular
    m["absorption"][0, :3] = absorption
    m["parameters"][0] = [pk(metallic, subsurface, specular_f, roughness), pk(specular_tint, anisotropic, sheen, sheen_tint),
                          pk(clearcoat, clearcoat_gloss, transmission, eta), 0]
    for k in ("diffuse_map", "normal_map", "metallic_roughness_map", "emissive_map", "sheen_map"):
        m[k] = -1
    # MaterialProps flag bits (crates/rfw-scene/src/material/mod.rs:27-34): bit 0 diffuse map, bit 1 normal map
    if diffuse_map >= 0:
        m["diffuse_map"] = diffuse_map; m["flags"] |= 1
    if normal_map >= 0:
        m["normal_map"] = normal_map; m["flags"] |= 2
    return m


class Texture:
    """TextureData as it crosses the boundary (crates/rfw-backend/src/structs.rs:197-249): all mip levels in one
    byte slice, level l at texel offset sum_{i<l} (w>>i)*(h>>i); format 0 = BGRA8, 1 = RGBA8."""

    def __init__(self, level0, mip_levels=5, fmt=0):
        img = np.ascontiguousarray(level0, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 4
        self.height, self.width = img.shape[:2]
        self.format = fmt
        levels = [img]
        while len(levels) < mip_levels and levels[-1].shape[0] > 1 and levels[-1].shape[1] > 1:
            a = levels[-1].astype(np.uint16)
            h2, w2 = a.shape[0] // 2, a.shape[1] // 2
            a = a[: 2 * h2, : 2 * w2]
            levels.append(((a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2] + 2) // 4).astype(np.uint8))  # 2x2 box filter
        self.mip_levels = len(levels)
        self.levels = levels
        self.bytes = np.concatenate([l.reshape(-1) for l in levels])


def pattern_texture(size=64, seed=SEED_SCENE, kind="checker", fmt=0, mip_levels=5):
    """Deterministic test textures: `checker` (coloured 8x8 checkerboard with a per-texel hash jitter), `normal`
    (a smooth bump field encoded as a tangent-space normal map), `sky` (equirect gradient with a bright sun disc)."""
    y, x = np.mgrid[0:size, 0: size * (2 if kind == "sky" else 1)].astype(np.float64)
    w = x.shape[1]
    jit = u01(seed, (y * w + x).astype(np.uint64)).astype(np.float64)
    if kind == "checker":
        c = ((x // (size // 8)).astype(int) + (y // (size // 8)).astype(int)) % 2
        r = np.where(c, 0.9, 0.25) * (0.8 + 0.2 * jit); g = np.where(c, 0.6, 0.3) * (0.8 + 0.2 * jit); b = np.where(c, 0.3, 0.8)
    elif kind == "normal":
        fx = 0.35 * np.sin(2 * np.pi * 3 * x / size); fy = 0.35 * np.cos(2 * np.pi * 2 * y / size)
        nz = 1.0 / np.sqrt(1 + fx * fx + fy * fy)
        r, g, b = 0.5 + 0.5 * fx * nz, 0.5 + 0.5 * fy * nz, 0.5 + 0.5 * nz
    else:
        v = y / size
        r = 0.25 + 0.5 * v; g = 0.35 + 0.4 * v; b = 0.9 - 0.3 * v
        sun = ((x / w - 0.3) ** 2 * 4 + (v - 0.7) ** 2) < 0.002
        r, g, b = np.where(sun, 1.0, r), np.where(sun, 1.0, g), np.where(sun, 0.9, b)
    rgba = np.stack([r, g, b, np.ones_like(r)], axis=2)
    img = np.clip(np.round(rgba * 255.0), 0, 255).astype(np.uint8)
    if fmt == 0:
        img = img[:, :, [2, 1, 0, 3]]  # stored as BGRA8
    return Texture(img, mip_levels, fmt)


def set_uvs(tris, scale=1.0, lod=0.0):
    """Planar UVs from the two dominant axes of each triangle's normal (what a loader's UV set provides)."""
    n = np.abs(tris["normal"])
    ax = np.argmax(n, axis=1)
    ua = np.where(ax == 0, 1, 0); va = np.where(ax == 2, 1, 2)
    idx = np.arange(len(tris))
    for k, vert in (("0", "vertex0"), ("1", "vertex1"), ("2", "vertex2")):
        tris["u" + k] = tris[vert][idx, ua] * scale
        tris["v" + k] = tris[vert][idx, va] * scale
    tris["lod"] = lod
    return tris


def area_lights_from(tris, matrix, radiance, mesh_id, inst_idx, first_light_id=0):
    """AreaLight::new for every triangle of an emissive instance (lights.rs:71-97, scene update_lights
    crates/rfw-scene/src/lib.rs:575-648) — world-space vertices; also returns the light_id per triangle."""
    M = np.asarray(matrix, np.float64).reshape(4, 4)

    def xf(p):
        return (p.astype(np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)

    v0, v1, v2 = xf(tris["vertex0"]), xf(tris["vertex1"]), xf(tris["vertex2"])
    nm = np.linalg.inv(M[:3, :3]).T
    nrm = _norm(tris["normal"].astype(np.float64) @ nm.T).astype(np.float32)
    L = np.zeros(len(tris), dtype=wire.AREA_LIGHT)
    L["position"] = (v0 + v1 + v2) * np.float32(1.0 / 3.0)
    rad = np.abs(np.asarray(radiance, np.float32))
    L["radiance"] = rad
    L["energy"] = np.float32(np.linalg.norm(rad))
    L["normal"] = nrm
    a = np.linalg.norm((v1 - v0).astype(np.float64), axis=1); b = np.linalg.norm((v2 - v1).astype(np.float64), axis=1); c = np.linalg.norm((v0 - v2).astype(np.float64), axis=1)
    s = (a + b + c) * 0.5
    L["area"] = np.sqrt(np.maximum(s * (s - a) * (s - b) * (s - c), 0)).astype(np.float32)
    L["vertex0"], L["vertex1"], L["vertex2"] = v0, v1, v2
    L["inst_idx"] = inst_idx; L["mesh_id"] = mesh_id; L["_dummy1"] = 1; L["_dummy2"] = 2
    ids = np.arange(first_light_id, first_light_id + len(tris), dtype=np.int32)
    return L, ids


def bounds_of(tris):
    allv = np.concatenate([tris["vertex0"], tris["vertex1"], tris["vertex2"]])
    return allv.min(axis=0), allv.max(axis=0)


# ---- whole-scene descriptions used by tests / bench --------------------------------------------------
class SceneDesc:
    """Plain description of what is sent through the boundary: meshes {id: RTTriangle[]}, instances
    {mesh id: (n,16) column-major}, materials, lights.  `apply(backend)` replays the call order of
    rfw's synchronize_system (rfw/src/system/mod.rs:19-206) on anything exposing the Backend methods
    (the CUDA backend and the oracle adapter alike)."""

    def __init__(self):
        self.meshes = {}
        self.instances = {}
        self.materials = np.zeros(0, dtype=wire.DEVICE_MATERIAL)
        self.area_lights = np.zeros(0, dtype=wire.AREA_LIGHT)
        self.point_lights = np.zeros(0, dtype=wire.POINT_LIGHT)
        self.spot_lights = np.zeros(0, dtype=wire.SPOT_LIGHT)
        self.directional_lights = np.zeros(0, dtype=wire.DIRECTIONAL_LIGHT)
        self.textures = []   # list of Texture
        self.skybox = None   # Texture or None
        self.skin_data = {}        # mesh id -> JOINT_DATA[3 * triangles] (MeshData3D::skin_data)
        self.instance_skins = {}   # mesh id -> int32 skin id per instance (-1 = none)
        self.skins = []            # list of (n_joints, 16) column-major joint matrices

    def apply(self, backend):
        for mid, tris in self.meshes.items():
            if mid in self.skin_data:
                backend.set_3d_mesh(mid, tris, skin_data=self.skin_data[mid])
            else:
                backend.set_3d_mesh(mid, tris)
        for mid, mats in self.instances.items():
            if mid in self.instance_skins:
                backend.set_3d_instances(mid, mats, skin_ids=self.instance_skins[mid])
            else:
                backend.set_3d_instances(mid, mats)
        if self.skins:
            backend.set_skins(self.skins)
        backend.set_materials(self.materials)
        if self.textures or self.skybox is not None:
            backend.set_textures(self.textures)
            backend.set_skybox(self.skybox)
        backend.set_area_lights(self.area_lights)
        backend.set_point_lights(self.point_lights)
        backend.set_spot_lights(self.spot_lights)
        backend.set_directional_lights(self.directional_lights)
        backend.synchronize()


def soup_scene(n, s, seed=SEED_SCENE):
    sc = SceneDesc()
    sc.meshes[0] = soup(n, s, seed)
    sc.instances[0] = to_column_major([identity()])
    sc.materials = material()
    return sc


def instanced_scene(grid=100, subdiv=3, seed=SEED_SCENE, n_lights=16, ground=True):
    """C3: grid x grid icosphere instances with random rotation / uniform scale in [0.5,1],
    a ground quad, n_lights emissive quads at y = 6 facing down, 8 materials alternating Lambert / GGX metal."""
    sc = SceneDesc()
    mats = []
    for i in range(8):
        col = (0.35 + 0.6 * u01(seed + 7, np.arange(i * 3, i * 3 + 3))).astype(np.float32)
        if i % 2 == 0:
            mats.append(material(color=col, metallic=0.0, roughness=1.0))
        else:
            mats.append(material(color=col, metallic=1.0, roughness=[0.2, 0.4, 0.6, 0.4][i // 2]))
    mats.append(material(color=(0.6, 0.6, 0.6), roughness=1.0))  # 8: ground
    mats.append(material(color=(12.0, 12.0, 12.0)))              # 9: emissive (colour > 1, material/list.rs:494)
    sc.materials = np.concatenate(mats)
    # one sphere mesh per material would need 8 meshes; the reference example uses one mesh and per-mesh
    # materials, so: 8 sphere meshes (mesh id m uses material m), instances dealt round-robin.
    for m in range(8):
        sc.meshes[m] = icosphere(subdiv, 0.4, mat_id=m)
    n = grid * grid
    k = np.arange(n)
    r = u01(seed + 1, k[:, None] * 8 + np.arange(8)[None, :]).astype(np.float64)
    per_mesh = {m: [] for m in range(8)}
    for i in range(n):
        gx, gz = i % grid, i // grid
        axis = _norm(np.array([r[i, 0] - 0.5, r[i, 1] - 0.5, r[i, 2] - 0.5]) + 1e-6)
        M = trs((gx - grid / 2 + 0.5, 0.4, gz - grid / 2 + 0.5), axis, r[i, 3] * 2 * np.pi, 0.5 + 0.5 * r[i, 4])
        per_mesh[i % 8].append(M)
    for m in range(8):
        sc.instances[m] = to_column_major(per_mesh[m]) if per_mesh[m] else np.zeros((0, 16), np.float32)
    inst_base = n
    next_mesh = 8
    if ground:
        g = quad((0, 0, 0), (0, 1, 0), grid * 1.2, grid * 1.2, mat_id=8)
        if g["normal"][0, 1] < 0:  # make the ground face up
            g = make_triangles(g["vertex0"], g["vertex2"], g["vertex1"], 8)
        sc.meshes[next_mesh] = g
        sc.instances[next_mesh] = to_column_major([identity()])
        next_mesh += 1
        inst_base += 1
    if n_lights:
        lq = quad((0, 0, 0), (0, 1, 0), 3.0, 3.0, mat_id=9)
        if lq["normal"][0, 1] > 0:  # face down
            lq = make_triangles(lq["vertex0"], lq["vertex2"], lq["vertex1"], 9)
        side = int(np.ceil(np.sqrt(n_lights)))
        lmats, lights = [], []
        for i in range(n_lights):
            lx = ((i % side) + 0.5) / side * grid - grid / 2
            lz = ((i // side) + 0.5) / side * grid - grid / 2
            lmats.append(trs((lx, 6.0, lz)))
        light_tris = lq.copy()
        # light_id is per triangle of the MESH, but area lights are per instance; the reference writes the
        # id of the last instance processed (crates/rfw-scene/src/lib.rs:640-645).  light_id only feeds
        # LightPickProb, which is uniform (shade.comp:368), so any valid index gives identical results.
        for i, M in enumerate(lmats):
            L, ids = area_lights_from(lq, M, (12.0, 12.0, 12.0), next_mesh, inst_base + i, first_light_id=2 * i)
            lights.append(L)
            light_tris["light_id"] = ids
        sc.meshes[next_mesh] = light_tris
        sc.instances[next_mesh] = to_column_major(lmats)
        sc.area_lights = np.concatenate(lights)
    return sc


def soup_with_lights(n, s, n_lights=256, seed=SEED_SCENE, light_area=1e-2, radius=2.0, extra_ground=False):
    """C4 / C5 flavour: soup + emissive triangles on a sphere of `radius` around the cube, facing inward."""
    sc = soup_scene(n, s, seed)
    sc.materials = np.concatenate([material(color=(0.7, 0.7, 0.7), roughness=1.0), material(color=(12.0, 12.0, 12.0))])
    k = np.arange(n_lights)
    r = u01(SEED_LIGHTS, k[:, None] * 4 + np.arange(4)[None, :]).astype(np.float64)
    z = 1.0 - 2.0 * r[:, 0]; phi = 2 * np.pi * r[:, 1]; rad = np.sqrt(np.maximum(0, 1 - z * z))
    d = np.stack([rad * np.cos(phi), rad * np.sin(phi), z], axis=1)
    c = 0.5 + radius * d
    nrm = -d
    ref = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = _norm(np.cross(nrm, ref)); ty = np.cross(nrm, tx)
    e = np.sqrt(2.0 * light_area)  # right triangle with legs e: area = e^2/2
    v0 = c - (tx + ty) * e / 3.0
    v1 = v0 + tx * e
    v2 = v0 + ty * e
    lt = make_triangles(v0.astype(np.float32), v1.astype(np.float32), v2.astype(np.float32), 1)
    flip = np.einsum("ij,ij->i", lt["normal"].astype(np.float64), nrm) < 0
    if flip.any():
        a, b = lt["vertex1"].copy(), lt["vertex2"].copy()
        a[flip], b[flip] = lt["vertex2"][flip], lt["vertex1"][flip]
        lt = make_triangles(lt["vertex0"], a, b, 1)
    L, ids = area_lights_from(lt, identity(), (12.0, 12.0, 12.0), 1, 1, 0)
    lt["light_id"] = ids
    sc.meshes[1] = lt
    sc.instances[1] = to_column_major([identity()])
    sc.area_lights = L
    return sc


def mixed_scale_scene(n_small=100000, n_long=300, n_big=30, s=0.008, seed=SEED_SCENE):
    """One mesh mixing triangle scales the way authored assets do (a room's walls around its furniture): a soup of small triangles,
    long thin needles crossing the whole cube and a few cube-sized triangles — the case spatial splits exist for."""
    base = soup(n_small, s, seed=seed)
    k = np.arange(n_long + n_big)
    r = u01(seed + 17, k[:, None] * 9 + np.arange(9)[None, :]).astype(np.float32)
    a = r[:, 0:3]
    b = r[:, 3:6]
    c = np.where((k < n_long)[:, None], a + (b - a) * 0.5 + (r[:, 6:9] - 0.5) * np.float32(0.01), r[:, 6:9])   # needles: third vertex near the middle of a-b
    extra = make_triangles(a, b, c.astype(np.float32), 0)
    sc = SceneDesc()
    sc.materials = material()
    sc.meshes[0] = np.concatenate([base, extra])
    sc.meshes[0]["id"] = np.arange(len(sc.meshes[0]), dtype=np.int32)
    sc.instances[0] = to_column_major([identity()])
    return sc


def random_barycentrics_np(r0):
    """Vectorised RandomBarycentrics (shade.comp:372-412): 16 steps of base-4 triangle subdivision driven by the bits of r0."""
    uf = (r0.astype(np.float64) * 4294967295.0).astype(np.uint64).astype(np.uint32)
    A = np.stack([np.ones_like(r0), np.zeros_like(r0)], 1).astype(np.float32)
    B = np.stack([np.zeros_like(r0), np.ones_like(r0)], 1).astype(np.float32)
    C = np.zeros_like(A)
    for i in range(16):
        d = ((uf >> np.uint32(2 * (15 - i))) & np.uint32(3))[:, None]
        An = np.where(d == 0, (B + C) * 0.5, np.where(d == 1, A, np.where(d == 2, (B + A) * 0.5, (C + A) * 0.5)))
        Bn = np.where(d == 0, (A + C) * 0.5, np.where(d == 1, (A + B) * 0.5, np.where(d == 2, B, (C + B) * 0.5)))
        Cn = np.where(d == 0, (A + B) * 0.5, np.where(d == 1, (A + C) * 0.5, np.where(d == 2, (B + C) * 0.5, C)))
        A, B, C = An.astype(np.float32), Bn.astype(np.float32), Cn.astype(np.float32)
    r = (A + B + C) * np.float32(0.3333333)
    return np.stack([r[:, 0], r[:, 1], 1 - r[:, 0] - r[:, 1]], 1)


def c4_scene(n_tris=5_000_000):
    """BASELINE.json configs[3]: soup of `n_tris` triangles (s = 0.003) + 256 emissive triangles (area 1e-2) on a
    radius-2 sphere around the cube, facing inward."""
    return soup_with_lights(n_tris, 0.003, n_lights=256, light_area=1e-2, radius=2.0)


def c4_shadow_rays(desc, rays, hits):
    """The C4 workload: for every closest hit on the soup pick light floor(r0 * n_lights), sample a point on it with
    RandomBarycentrics (r0 re-scaled as shade.comp:474-475 does) and build the next-event-estimation any-hit ray
    t in (1e-3, dist - 2e-4) (ray_shadow.comp:254-257, shade.comp:253).  Returns (shadow rays, indices of the shading rays)."""
    from . import wire

    ok = hits["inst"] == 0  # shading points on the soup (not on the lights)
    idx = np.nonzero(ok)[0]
    P = rays["origin"][ok] + rays["direction"][ok] * hits["t"][ok][:, None]
    r0 = u01(SEED_LIGHTS + 1, idx)
    L = desc.area_lights
    li = np.minimum((r0 * len(L)).astype(np.int64), len(L) - 1)
    rb = (r0 - li.astype(np.float32) / np.float32(len(L))) * np.float32(len(L))
    bary = random_barycentrics_np(np.clip(rb, 0, 1).astype(np.float32))
    Q = L["vertex0"][li] * bary[:, :1] + L["vertex1"][li] * bary[:, 1:2] + L["vertex2"][li] * bary[:, 2:3]
    D = Q - P
    dist = np.linalg.norm(D, axis=1).astype(np.float32)
    sh = np.zeros(len(P), dtype=wire.RAY)
    sh["origin"] = P; sh["direction"] = D / dist[:, None]; sh["tmin"] = 1e-3; sh["tmax"] = dist - np.float32(2e-4)
    return sh, idx


def c5_scene(n_tris=10_000_000):
    """BASELINE.json configs[4]: soup of `n_tris` triangles (s = 0.002) + ground quad + 64 area-light triangles;
    rendered at 3840x2160, 64 spp, depth 5 from `c5_view`."""
    desc = soup_with_lights(n_tris, 0.002, n_lights=64, light_area=2e-2, radius=2.0)
    g = quad((0.5, -0.2, 0.5), (0, 1, 0), 6.0, 6.0, mat_id=0)
    if g["normal"][0, 1] < 0:
        g = make_triangles(g["vertex0"], g["vertex2"], g["vertex1"], 0)
    desc.meshes[2] = g
    desc.instances[2] = to_column_major([identity()])
    return desc


def c5_view(w=3840, h=2160):
    return camera_view((0.5, 0.9, -2.2), (0.0, -0.15, 1.0), w, h)


def textured_scene(grid=4, subdiv=2, seed=SEED_SCENE, tex_size=64, skybox=True):
    """(f)1 flavour: a ground quad with a diffuse map + normal map, spheres with a diffuse map (one BGRA8, one RGBA8
    texture), plain GGX spheres, two emissive quads and an equirect skybox."""
    sc = SceneDesc()
    sc.textures = [pattern_texture(tex_size, seed, "checker", fmt=0), pattern_texture(tex_size, seed + 1, "normal", fmt=0),
                   pattern_texture(tex_size // 2, seed + 2, "checker", fmt=1)]
    if skybox:
        sc.skybox = pattern_texture(tex_size, seed + 3, "sky", fmt=0)
    sc.materials = np.concatenate([
        material(color=(0.9, 0.9, 0.9), roughness=1.0, diffuse_map=0, normal_map=1),   # 0 ground: diffuse + normal map
        material(color=(1.0, 0.95, 0.9), roughness=0.8, diffuse_map=2),                 # 1 textured spheres
        material(color=(0.8, 0.7, 0.5), metallic=1.0, roughness=0.3),                   # 2 metal spheres
        material(color=(12.0, 12.0, 12.0)),                                             # 3 emissive
    ])
    g = quad((0, 0, 0), (0, 1, 0), grid * 2.0, grid * 2.0, mat_id=0)
    if g["normal"][0, 1] < 0:
        g = make_triangles(g["vertex0"], g["vertex2"], g["vertex1"], 0)
    sc.meshes[0] = set_uvs(g, scale=0.37, lod=1.0)
    sc.instances[0] = to_column_major([identity()])
    sc.meshes[1] = set_uvs(icosphere(subdiv, 0.4, mat_id=1), scale=1.3, lod=4.0)
    sc.meshes[2] = icosphere(subdiv, 0.4, mat_id=2)
    per = {1: [], 2: []}
    r = u01(seed + 5, np.arange(grid * grid * 4)).astype(np.float64).reshape(-1, 4)
    for i in range(grid * grid):
        gx, gz = i % grid, i // grid
        axis = _norm(np.array([r[i, 0] - 0.5, r[i, 1] - 0.5, r[i, 2] - 0.5]) + 1e-6)
        per[1 + i % 2].append(trs((gx - grid / 2 + 0.5, 0.4, gz - grid / 2 + 0.5), axis, r[i, 3] * 2 * np.pi, 0.7 + 0.3 * r[i, 3]))
    for m in (1, 2):
        sc.instances[m] = to_column_major(per[m])
    lq = quad((0, 0, 0), (0, 1, 0), 1.5, 1.5, mat_id=3)
    if lq["normal"][0, 1] > 0:
        lq = make_triangles(lq["vertex0"], lq["vertex2"], lq["vertex1"], 3)
    lmats = [trs((-1.0, 3.0, 0.0)), trs((1.5, 3.5, 1.0))]
    inst_base = 1 + grid * grid
    lights = []
    lt = lq.copy()
    for i, M in enumerate(lmats):
        L, ids = area_lights_from(lq, M, (12.0, 12.0, 12.0), 3, inst_base + i, first_light_id=2 * i)
        lights.append(L); lt["light_id"] = ids
    sc.meshes[3] = lt
    sc.instances[3] = to_column_major(lmats)
    sc.area_lights = np.concatenate(lights)
    return sc


# ---- punctual lights (crates/rfw-backend/src/lights.rs:100-108, 199-209, 293-301) ----------------------
def _energy(radiance):
    return float(np.linalg.norm(np.asarray(radiance, np.float64)))  # `energy = radiance.length()`, as AreaLight::new does (lights.rs:71-97)


def point_light(position, radiance):
    l = np.zeros(1, dtype=wire.POINT_LIGHT)
    l["position"] = position; l["radiance"] = radiance; l["energy"] = _energy(radiance)
    return l


def spot_light(position, direction, inner_deg, outer_deg, radiance):
    """SpotLight::new (lights.rs:239-259): angles in degrees, cosines of the angles as given, |radiance|, normalised direction."""
    l = np.zeros(1, dtype=wire.SPOT_LIGHT)
    l["position"] = position; l["direction"] = _norm(np.asarray(direction, np.float64)).astype(np.float32)
    assert outer_deg > inner_deg
    l["cos_inner"] = np.cos(np.radians(inner_deg)); l["cos_outer"] = np.cos(np.radians(outer_deg))
    l["radiance"] = np.abs(radiance); l["energy"] = _energy(radiance)
    return l


def directional_light(direction, radiance):
    l = np.zeros(1, dtype=wire.DIRECTIONAL_LIGHT)
    l["direction"] = _norm(np.asarray(direction, np.float64)).astype(np.float32); l["radiance"] = radiance; l["energy"] = _energy(radiance)
    return l


def lights_and_lobes_scene(grid=6, subdiv=2, seed=SEED_SCENE, area_lights=2):
    """Every light type of RandomPointOnLight (shade.comp:414-528: area, point, spot, directional, in that index order)
    over instanced spheres whose 8 materials walk through the lobes of the Disney BSDF (disney.glsl:110-266): diffuse,
    subsurface, GGX metal, tinted dielectric specular, clearcoat, rough and smooth transmission (with absorption)."""
    sc = instanced_scene(grid=grid, subdiv=subdiv, seed=seed, n_lights=area_lights)
    mats = [
        material(color=(0.8, 0.3, 0.25), roughness=1.0),
        material(color=(0.3, 0.7, 0.4), roughness=0.6, subsurface=0.8),
        material(color=(0.9, 0.8, 0.5), metallic=1.0, roughness=0.3),
        material(color=(0.3, 0.4, 0.9), roughness=0.35, specular_f=0.9, specular_tint=0.7),
        material(color=(0.7, 0.2, 0.6), roughness=0.5, clearcoat=1.0, clearcoat_gloss=0.9),
        material(color=(0.95, 0.95, 0.95), roughness=0.25, transmission=0.9, eta=0.66, absorption=(0.4, 0.1, 0.05)),
        material(color=(0.9, 0.95, 1.0), roughness=0.05, transmission=1.0, eta=0.75),
        material(color=(0.6, 0.6, 0.3), metallic=0.5, roughness=0.15, clearcoat=0.5, clearcoat_gloss=0.3, transmission=0.3, eta=0.8),
    ]
    sc.materials = np.concatenate(mats + [sc.materials[8:9], sc.materials[9:10]])
    h = grid / 2
    sc.point_lights = np.concatenate([point_light((-0.6 * h, 2.5, -0.5 * h), (30.0, 24.0, 18.0)), point_light((0.7 * h, 1.5, 0.4 * h), (10.0, 16.0, 28.0))])
    sc.spot_lights = spot_light((0.0, 5.0, -0.8 * h), (0.1, -1.0, 0.35), 25.0, 50.0, (90.0, 90.0, 80.0))
    sc.directional_lights = directional_light((0.4, -1.0, 0.3), (1.2, 1.1, 1.0))
    return sc
