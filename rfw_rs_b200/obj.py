"""Minimal Wavefront OBJ / MTL reader for config C1 ("one obj/gltf mesh from assets/models").

Plays the role of the reference's loader glue (crates/rfw-scene/src/loaders/obj.rs:26-255, which delegates the parsing to the
third-party `tobj` crate with single_index + triangulate + ignore_points + ignore_lines) and follows what that glue does with
the parsed data:
  * ALL objects / groups of the file end up in ONE Mesh3D, as a flat triangle list (obj.rs:197-253);
  * polygons are fan-triangulated around their first vertex (tobj's triangulate);
  * a face takes the material named by the last `usemtl` (tobj splits models there); faces without one take material 0,
    and a file without any material gets ONE red fallback material (obj.rs:186-193);
  * per material (obj.rs:48-183): colour = Kd, raised to the emission Ke where that is larger — an emission with every
    component <= 1 is multiplied by 10 first (obj.rs:84-100); roughness = clamp(1 - log10(Ns) / 1000, 0, 1) (obj.rs:52-54;
    Ns = 0 gives -inf -> +inf -> 1); transmission = 1 - d (obj.rs:55); eta = Ni (obj.rs:56); specular = Ks;
  * missing normals are passed on as zeros and Mesh3D::new derives flat normals (objects_3d/mod.rs:331-383) — here
    scenes.make_triangles does; missing texture coordinates are zeros.
Texture maps (map_Kd, norm, map_Pr ...) are returned by name in `ObjAsset.texture_names` but not decoded (no image decoder in
the harness; textures reach the backend as RGBA arrays through set_textures, tests/test_textures.py).

The backend only ever sees the RTTriangle / DeviceMaterial arrays this produces.
"""
import math
import os

import numpy as np

from . import scenes, wire


class ObjAsset:
    def __init__(self):
        self.positions = np.zeros((0, 3, 3), np.float32)   # (triangles, corner, xyz)
        self.normals = None                                # (triangles, corner, xyz) or None
        self.uvs = None                                    # (triangles, corner, uv) or None
        self.material_ids = np.zeros(0, np.int32)          # per triangle, index into `materials`
        self.materials = np.zeros(0, dtype=wire.DEVICE_MATERIAL)
        self.material_names = []
        self.texture_names = {}                            # material name -> {key: file name}


def _floats(tok, n, default=0.0):
    out = [default] * n
    for i, t in enumerate(tok[:n]):
        try:
            out[i] = float(t)
        except ValueError:
            out[i] = 0.0
    return out


def parse_mtl(text):
    """-> (names, DeviceMaterial array, texture names).  Unknown keys are ignored like tobj's unknown_param ones the glue does not read."""
    mats, names, textures = [], [], {}
    cur = None

    def flush():
        if cur is None:
            return
        color = np.array(cur["kd"], np.float64)
        ke = np.array(cur["ke"], np.float64)
        if cur["has_ke"]:
            if not np.all(ke == 0.0) and np.all(ke <= 1.0):  # obj.rs:95-97
                ke = ke * 10.0
            color = np.maximum(ke, color)                    # obj.rs:99
        ns = cur["ns"]
        lg = -math.inf if ns == 0.0 else (math.nan if ns < 0.0 else math.log10(ns))
        rough = 1.0 - lg / 1000.0
        rough = 0.0 if math.isnan(rough) else min(max(rough, 0.0), 1.0)   # f32::max / min drop a NaN operand
        m = scenes.material(color=tuple(color), roughness=rough, specular=tuple(cur["ks"]), transmission=1.0 - cur["d"], eta=cur["ni"])
        mats.append(m)
        names.append(cur["name"])
        textures[cur["name"]] = cur["maps"]

    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        tok = line.split()
        key, args = tok[0].lower(), tok[1:]
        if key == "newmtl":
            flush()
            cur = {"name": " ".join(args), "kd": [0.0, 0.0, 0.0], "ks": [0.0, 0.0, 0.0], "ke": [0.0, 0.0, 0.0], "has_ke": False, "ns": 0.0, "d": 1.0, "ni": 1.0, "maps": {}}
            # (tobj defaults: diffuse / specular 0, shininess 0, dissolve 1, optical_density 1)
        elif cur is None:
            continue
        elif key == "kd":
            cur["kd"] = _floats(args, 3)
        elif key == "ks":
            cur["ks"] = _floats(args, 3)
        elif key == "ke":
            cur["ke"] = _floats(args, 3); cur["has_ke"] = True
        elif key == "ns":
            cur["ns"] = _floats(args, 1)[0]
        elif key == "d":
            cur["d"] = _floats(args, 1, 1.0)[0]
        elif key == "tr":
            cur["d"] = 1.0 - _floats(args, 1)[0]
        elif key == "ni":
            cur["ni"] = _floats(args, 1, 1.0)[0]
        elif key in ("map_kd", "map_bump", "bump", "norm", "map_ns", "map_pr", "map_pm", "pm", "map_ke", "map_ps", "ps") and args:
            cur["maps"][key] = args[-1]
    flush()
    arr = np.concatenate(mats) if mats else np.zeros(0, dtype=wire.DEVICE_MATERIAL)
    return names, arr, textures


def _index(tok, count):
    """OBJ indices are 1-based; negative ones count back from the end of what has been read so far."""
    i = int(tok)
    return i - 1 if i > 0 else count + i


def parse_obj(text, mtl_loader=None):
    """`mtl_loader(name) -> text or None` resolves `mtllib` lines (None: the file has no readable material library)."""
    pos, nrm, tex = [], [], []
    tri_p, tri_n, tri_t, tri_m = [], [], [], []
    names, mats, textures = [], np.zeros(0, dtype=wire.DEVICE_MATERIAL), {}
    cur_mat = -1
    any_n = any_t = False
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        tok = line.split()
        key, args = tok[0], tok[1:]
        if key == "v":
            pos.append(_floats(args, 3))
        elif key == "vn":
            nrm.append(_floats(args, 3))
        elif key == "vt":
            tex.append(_floats(args, 2))
        elif key == "mtllib" and mtl_loader is not None:
            lib = mtl_loader(" ".join(args))
            if lib is not None:
                n2, m2, t2 = parse_mtl(lib)
                names += n2
                mats = np.concatenate([mats, m2])
                textures.update(t2)
        elif key == "usemtl":
            name = " ".join(args)
            cur_mat = names.index(name) if name in names else -1
        elif key == "f" and len(args) >= 3:   # points ("p") and lines ("l") are ignored (ignore_points / ignore_lines)
            corners = []
            for a in args:
                parts = a.split("/")
                vi = _index(parts[0], len(pos))
                ti = _index(parts[1], len(tex)) if len(parts) > 1 and parts[1] else None
                ni = _index(parts[2], len(nrm)) if len(parts) > 2 and parts[2] else None
                corners.append((vi, ti, ni))
            for k in range(1, len(corners) - 1):   # fan
                c3 = (corners[0], corners[k], corners[k + 1])
                tri_p.append([pos[c[0]] for c in c3])
                tri_n.append([nrm[c[2]] if c[2] is not None else [0.0, 0.0, 0.0] for c in c3])
                tri_t.append([tex[c[1]] if c[1] is not None else [0.0, 0.0] for c in c3])
                any_n = any_n or all(c[2] is not None for c in c3)
                any_t = any_t or all(c[1] is not None for c in c3)
                tri_m.append(max(cur_mat, 0))   # no / unknown material: the first one (obj.rs:232-240)
    out = ObjAsset()
    out.positions = np.array(tri_p, np.float32).reshape(-1, 3, 3)
    out.normals = np.array(tri_n, np.float32).reshape(-1, 3, 3) if any_n else None
    out.uvs = np.array(tri_t, np.float32).reshape(-1, 3, 2) if any_t else None
    out.material_ids = np.array(tri_m, np.int32)
    if len(mats) == 0:   # obj.rs:186-193: one red material
        mats = scenes.material(color=(1.0, 0.0, 0.0), roughness=1.0, specular=(0.0, 0.0, 0.0), transmission=1.0)
        names = ["<fallback>"]
    out.materials, out.material_names, out.texture_names = mats, names, textures
    return out


def load(path):
    base = os.path.dirname(path)

    def mtl(name):
        p = os.path.join(base, name)
        return open(p).read() if os.path.exists(p) else None

    return parse_obj(open(path).read(), mtl)


def triangles(asset):
    """RTTriangle records of the whole file (one Mesh3D).  Corners without a normal in the file get the flat normal; zero-area
    faces are dropped (their normal is NaN in the reference)."""
    p = asset.positions
    cr = np.cross((p[:, 1] - p[:, 0]).astype(np.float64), (p[:, 2] - p[:, 0]).astype(np.float64))
    ln = np.linalg.norm(cr, axis=1)
    keep = ln > 0
    p, cr, ln = p[keep], cr[keep], ln[keep]
    kw = {}
    if asset.normals is not None:
        n = asset.normals[keep].astype(np.float64)
        l2 = np.linalg.norm(n, axis=2, keepdims=True)
        flat = cr / ln[:, None]
        n = np.where(l2 > 0, n / np.maximum(l2, 1e-300), flat[:, None, :]).astype(np.float32)
        kw = {"n0": n[:, 0], "n1": n[:, 1], "n2": n[:, 2]}
    t = scenes.make_triangles(p[:, 0], p[:, 1], p[:, 2], mat_id=asset.material_ids[keep], **kw)
    if asset.uvs is not None:
        uv = asset.uvs[keep]
        for k in range(3):   # RTTriangle carries the texture coordinates in the w slots of its first six vectors (structs.rs:905-960)
            t[f"u{k}"] = uv[:, k, 0]
            t[f"v{k}"] = uv[:, k, 1]
    return t


def scene(asset):
    """One mesh, identity instance, the file's materials.  Triangles of an emissive material (a colour component above 1,
    material/mod.rs:79-84) become area lights the way rfw's light extraction does (scenes.area_lights_from)."""
    sc = scenes.SceneDesc()
    t = triangles(asset)
    sc.materials = asset.materials
    emissive = np.array([bool(np.any(m["color"][:3] > 1.0)) for m in asset.materials], bool)
    is_light = emissive[t["mat_id"]] if len(t) else np.zeros(0, bool)
    if is_light.any():
        lights = []
        idx = np.nonzero(is_light)[0]
        for j, k in enumerate(idx):
            L, ids = scenes.area_lights_from(t[k:k + 1], scenes.identity(), asset.materials[t["mat_id"][k]]["color"][:3], 0, 0, first_light_id=j)
            t["light_id"][k] = ids[0]
            lights.append(L)
        sc.area_lights = np.concatenate(lights)
    sc.meshes[0] = t
    sc.instances[0] = scenes.to_column_major([scenes.identity()])
    return sc
