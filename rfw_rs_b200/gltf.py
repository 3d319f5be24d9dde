"""Minimal glTF 2.0 reader for config C1 (primary-ray casting of an asset mesh).

Plays the role of the reference's loader glue (crates/rfw-scene/src/loaders/gltf.rs:27-100, which delegates the
parsing to the third-party `l3d` crate) for plain glTF: external .bin buffers, float32 POSITION / NORMAL, u16 / u32
indices, triangle lists, node matrices or TRS.  Per-vertex JOINTS_0 / WEIGHTS_0 and the skins' inverse bind matrices
are read too (the payload of MeshData3D::skin_data / SkinData); animation channels are not evaluated —
the backend only ever sees final joint matrices, which the tests synthesise (pose_joints).  Two ways to hand the asset to the backend:
  * flatten(asset)  -> one mesh, node transforms baked in, identity instance                       (variant C1a)
  * per_mesh(asset) -> one mesh per glTF mesh, one instance per mesh-bearing node with its matrix  (variant C1b:
                       what rfw really does: loaders/gltf.rs:73-79, graph/mod.rs:435-445)
Both must return the same world-space hits.
"""
import json
import os

import numpy as np

from . import scenes

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class Asset:
    def __init__(self):
        self.meshes = []      # list of dict(positions (n,3) f32, normals (n,3) f32 or None, indices (m,3) u32)
        self.mesh_nodes = []  # list of (mesh index, 4x4 float64 world matrix, row-indexed)
        self.skins = []       # list of (n_joints, 4, 4) float64 inverse bind matrices (row-indexed)
        self.mesh_skin = {}   # mesh index -> skin index (from the node that instantiates it)


def _accessor(g, buffers, idx):
    a = g["accessors"][idx]
    bv = g["bufferViews"][a["bufferView"]]
    dt = np.dtype(_COMP[a["componentType"]])
    nc = _NCOMP[a["type"]]
    off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    stride = bv.get("byteStride", 0)
    buf = buffers[bv["buffer"]]
    if stride and stride != dt.itemsize * nc:
        arr = np.lib.stride_tricks.as_strided(np.frombuffer(buf, dtype=dt, offset=off, count=1), shape=(a["count"], nc), strides=(stride, dt.itemsize))
        return np.array(arr)
    return np.frombuffer(buf, dtype=dt, offset=off, count=a["count"] * nc).reshape(a["count"], nc).copy()


def _node_matrix(n):
    if "matrix" in n:
        return np.array(n["matrix"], np.float64).reshape(4, 4).T  # glTF stores column-major
    M = np.eye(4)
    if "scale" in n:
        M = np.diag(list(n["scale"]) + [1.0]) @ M
    if "rotation" in n:
        x, y, z, w = n["rotation"]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0],
                      [0, 0, 0, 1]], np.float64)
        M = R @ M
    if "translation" in n:
        T = np.eye(4)
        T[:3, 3] = n["translation"]
        M = T @ M
    return M


def load(path):
    g = json.load(open(path))
    base = os.path.dirname(path)
    buffers = [open(os.path.join(base, b["uri"]), "rb").read() for b in g["buffers"]]
    asset = Asset()
    for m in g["meshes"]:
        pos, nrm, idx, jnt, wgt = [], [], [], [], []
        voff = 0
        for pr in m["primitives"]:
            if pr.get("mode", 4) != 4:
                continue
            p = _accessor(g, buffers, pr["attributes"]["POSITION"]).astype(np.float32)
            n = _accessor(g, buffers, pr["attributes"]["NORMAL"]).astype(np.float32) if "NORMAL" in pr["attributes"] else np.zeros_like(p)
            if "indices" in pr:
                i = _accessor(g, buffers, pr["indices"]).astype(np.uint32).reshape(-1, 3)
            else:
                i = np.arange(len(p), dtype=np.uint32).reshape(-1, 3)
            pos.append(p); nrm.append(n); idx.append(i + voff)
            if "JOINTS_0" in pr["attributes"] and "WEIGHTS_0" in pr["attributes"]:
                jnt.append(_accessor(g, buffers, pr["attributes"]["JOINTS_0"]).astype(np.uint32))
                w = _accessor(g, buffers, pr["attributes"]["WEIGHTS_0"])
                wgt.append(w.astype(np.float32) / (1.0 if w.dtype == np.float32 else float(np.iinfo(w.dtype).max)))
            voff += len(p)
        mesh = {"positions": np.concatenate(pos), "normals": np.concatenate(nrm), "indices": np.concatenate(idx)}
        if jnt and sum(len(j) for j in jnt) == len(mesh["positions"]):
            mesh["joints"], mesh["weights"] = np.concatenate(jnt), np.concatenate(wgt)
        asset.meshes.append(mesh)
    for sk in g.get("skins", []):
        ibm = _accessor(g, buffers, sk["inverseBindMatrices"]).astype(np.float64).reshape(-1, 4, 4).transpose(0, 2, 1) if "inverseBindMatrices" in sk else np.tile(np.eye(4), (len(sk["joints"]), 1, 1))
        asset.skins.append(ibm)
    scene = g["scenes"][g.get("scene", 0)]

    def walk(ni, parent):
        n = g["nodes"][ni]
        M = parent @ _node_matrix(n)
        if "mesh" in n:
            asset.mesh_nodes.append((n["mesh"], M))
            if "skin" in n:
                asset.mesh_skin[n["mesh"]] = n["skin"]
        for c in n.get("children", []):
            walk(c, M)

    for r in scene["nodes"]:
        walk(r, np.eye(4))
    return asset


def save_npz(asset, path):
    """Compact fixture (tests/golden): concatenated vertex/index arrays + offsets + node matrices."""
    pos = np.concatenate([m["positions"] for m in asset.meshes])
    nrm = np.concatenate([m["normals"] for m in asset.meshes]).astype(np.float16)
    idx = np.concatenate([m["indices"] for m in asset.meshes])
    voffs = np.cumsum([0] + [len(m["positions"]) for m in asset.meshes]).astype(np.int64)
    ioffs = np.cumsum([0] + [len(m["indices"]) for m in asset.meshes]).astype(np.int64)
    nodes_mesh = np.array([mi for mi, _ in asset.mesh_nodes], np.int32)
    nodes_mat = np.array([M for _, M in asset.mesh_nodes], np.float64)
    extra = {}
    if asset.skins and all("joints" in m for m in asset.meshes):
        extra["joints"] = np.concatenate([m["joints"] for m in asset.meshes]).astype(np.uint8 if max(len(s) for s in asset.skins) < 256 else np.uint16)
        extra["weights"] = np.concatenate([m["weights"] for m in asset.meshes]).astype(np.float16)
        extra["skin0_ibm"] = asset.skins[0].astype(np.float32)
        extra["mesh_skin"] = np.array([asset.mesh_skin.get(k, -1) for k in range(len(asset.meshes))], np.int32)
    np.savez_compressed(path, pos=pos, nrm=nrm, idx=idx if idx.max() > 65535 else idx.astype(np.uint16), voffs=voffs, ioffs=ioffs, nodes_mesh=nodes_mesh, nodes_mat=nodes_mat, **extra)


def load_npz(path):
    z = np.load(path)
    asset = Asset()
    for k in range(len(z["voffs"]) - 1):
        v0, v1 = z["voffs"][k], z["voffs"][k + 1]
        i0, i1 = z["ioffs"][k], z["ioffs"][k + 1]
        mesh = {"positions": z["pos"][v0:v1].astype(np.float32), "normals": z["nrm"][v0:v1].astype(np.float32), "indices": z["idx"][i0:i1].astype(np.uint32)}
        if "joints" in z.files:
            w = z["weights"][v0:v1].astype(np.float32)
            mesh["joints"], mesh["weights"] = z["joints"][v0:v1].astype(np.uint32), w / np.maximum(w.sum(axis=1, keepdims=True), 1e-8)  # f16 storage: renormalise
        asset.meshes.append(mesh)
    asset.mesh_nodes = [(int(mi), M) for mi, M in zip(z["nodes_mesh"], z["nodes_mat"])]
    if "skin0_ibm" in z.files:
        asset.skins = [z["skin0_ibm"].astype(np.float64)]
        asset.mesh_skin = {k: int(v) for k, v in enumerate(z["mesh_skin"]) if v >= 0}
    return asset


def joint_data(mesh):
    """MeshData3D::skin_data for the triangles _tris() builds: one JointData per triangle vertex, 3 per triangle."""
    from . import wire
    i = mesh["indices"].astype(np.int64).reshape(-1)
    jd = np.zeros(len(i), dtype=wire.JOINT_DATA)
    jd["joint"], jd["weight"] = mesh["joints"][i], mesh["weights"][i]
    return jd


def pose_joints(ibm, angle=0.35, seed=3):
    """Synthetic pose: every joint rotates about its own bind position by a seeded axis/angle, composed down a chain of
    decreasing strength — final joint matrices (world-from-bind), column-major float32 (n, 16).  Stands in for the
    animation evaluation of the scene graph (crates/rfw-scene/src/graph), which is outside the backend."""
    n = len(ibm)
    out = np.zeros((n, 16), np.float32)
    r = scenes.u01(seed, np.arange(n * 4)).reshape(n, 4).astype(np.float64)
    for j in range(n):
        bind = np.linalg.inv(ibm[j])           # joint's bind-pose world matrix
        c = bind[:3, 3]
        axis = r[j, :3] - 0.5
        axis /= max(np.linalg.norm(axis), 1e-9)
        a = angle * (2.0 * r[j, 3] - 1.0)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = c - R @ c                    # rotate about the joint's bind position
        out[j] = M.T.reshape(-1)                # column-major
    return out


def _tris(mesh, M=None, mat_id=0):
    p, n, i = mesh["positions"], mesh["normals"], mesh["indices"].astype(np.int64)
    if M is not None:
        p = (p.astype(np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
        nm = np.linalg.inv(M[:3, :3]).T
        n = n.astype(np.float64) @ nm.T
        n = (n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)).astype(np.float32)
    have_n = np.abs(n).sum() > 0
    return scenes.make_triangles(p[i[:, 0]], p[i[:, 1]], p[i[:, 2]], mat_id, *( (n[i[:, 0]], n[i[:, 1]], n[i[:, 2]]) if have_n else (None, None, None)))


def flatten(asset):
    """Variant C1a: one mesh with the node transforms baked in, identity instance."""
    parts = [_tris(asset.meshes[mi], M) for mi, M in asset.mesh_nodes]
    sc = scenes.SceneDesc()
    sc.meshes[0] = scenes.concat_meshes(parts)
    sc.instances[0] = scenes.to_column_major([scenes.identity()])
    sc.materials = scenes.material()
    return sc


def per_mesh(asset):
    """Variant C1b: one mesh per glTF mesh, one instance per mesh-bearing node (BLAS per mesh + TLAS)."""
    sc = scenes.SceneDesc()
    inst = {}
    for mi, M in asset.mesh_nodes:
        inst.setdefault(mi, []).append(M.astype(np.float32))
    for mi, mats in inst.items():
        sc.meshes[mi] = _tris(asset.meshes[mi])
        sc.instances[mi] = scenes.to_column_major(mats)
    sc.materials = scenes.material()
    return sc


def c1_camera(sc, width=1280, height=720):
    """Camera of SURVEY §8d C1: Camera3D::new() defaults (fov 40 deg), aspect w/h, at centre + (0, 0.15 ext_y, -1.2 |ext|)
    looking at the AABB centre of the (world-space) scene."""
    los, his = [], []
    for mid, tris in sc.meshes.items():
        mats = np.asarray(sc.instances[mid], np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)
        lo, hi = scenes.bounds_of(tris)
        corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        for M in mats:
            w = corners @ M.T
            los.append(w[:, :3].min(axis=0)); his.append(w[:, :3].max(axis=0))
    lo, hi = np.min(los, axis=0), np.max(his, axis=0)
    centre, ext = (lo + hi) * 0.5, hi - lo
    pos = centre + np.array([0.0, 0.15 * ext[1], -1.2 * np.linalg.norm(ext)])
    return scenes.camera_view(pos, centre - pos, width, height, fov_deg=40.0)


def skinned(asset, pose=None, copies=1, spacing=1.5):
    """Skinned variant: per-mesh BLAS, `copies` instances of every skinned mesh side by side, each with
    skin id 0 except the last one of several (bind pose, skin id -1); joint matrices from pose_joints() unless given."""
    sc = per_mesh(asset)
    for mi, mesh in enumerate(asset.meshes):
        if "joints" not in mesh or mi not in sc.meshes:
            continue
        jd = joint_data(mesh)
        assert len(jd) == 3 * len(sc.meshes[mi]), "degenerate triangles were dropped: joint data no longer lines up"
        sc.skin_data[mi] = jd
        base = np.asarray(sc.instances[mi], np.float32).reshape(-1, 16)[0].reshape(4, 4).copy()
        mats, ids = [], []
        for c in range(copies):
            M = base.copy()
            M[3, 0] += spacing * c  # column-major: translation lives in the last row of the reshaped array
            mats.append(M.reshape(-1))
            ids.append(0 if (copies == 1 or c < copies - 1) else -1)
        sc.instances[mi] = np.array(mats, np.float32)
        sc.instance_skins[mi] = np.array(ids, np.int32)
    sc.skins = [pose_joints(asset.skins[0]) if pose is None else np.asarray(pose, np.float32)]
    return sc
