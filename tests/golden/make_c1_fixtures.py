#!/usr/bin/env python
"""Generates the C1 fixtures under tests/golden/ (run HERE, where /root/reference exists; the GPU box only reads the
committed .npz files):
  cesium_man.npz, pica.npz     geometry of assets/models/{CesiumMan/CesiumMan.gltf, pica/scene.gltf} (positions f32,
                               normals f16, indices, node matrices) — data files of the reference, not source code
  c1_<asset>_golden.npz        ORACLE closest hits (inst, prim, t) for the 1280x720 pinhole camera of SURVEY §8d C1 on a
                               4x4-subsampled pixel grid, variant C1a (flattened, det eps 0), plus the hit counts at the
                               reference's determinant epsilons 1e-4 (GLSL) and 1e-6 (Rust twin)
These are golden vectors of the ORACLE (the reference has none and cannot be run): a regression pin, not a reference pin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from rfw_rs_b200 import gltf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = {"cesium_man": "/root/reference/assets/models/CesiumMan/CesiumMan.gltf", "pica": "/root/reference/assets/models/pica/scene.gltf"}
W, H, SUB = 1280, 720, 4

for name, path in ASSETS.items():
    asset = gltf.load(path)
    gltf.save_npz(asset, os.path.join(HERE, name + ".npz"))
    asset = gltf.load_npz(os.path.join(HERE, name + ".npz"))  # the fixture is what the tests will see (f16 normals)
    sc = gltf.flatten(asset)
    view = gltf.c1_camera(sc, W, H)
    out = {}
    for eps in (0.0, 1e-6, 1e-4):
        o = orc.OracleBackend(det_eps=eps)
        sc.apply(o)
        rays = o.primary_rays(view, W, H).reshape(H, W)[::SUB, ::SUB].reshape(-1)
        hits = o.trace_closest(rays, mode=orc.MODE_BVH2)
        brute_ok = True
        if eps == 0.0:
            sel = np.arange(0, len(rays), 97)
            b = o.trace_closest(rays[sel], mode=orc.MODE_BRUTE)
            brute_ok = bool(np.array_equal(b["prim"], hits["prim"][sel]) and np.array_equal(b["t"], hits["t"][sel]))
            out.update(inst=hits["inst"].astype(np.int16), prim=hits["prim"], t=hits["t"], u=hits["u"].astype(np.float16), v=hits["v"].astype(np.float16))
        out[f"hit_count_eps_{eps:g}"] = np.int64((hits["inst"] >= 0).sum())
        print(name, "tris", len(sc.meshes[0]), "eps", eps, "hits", int((hits["inst"] >= 0).sum()), "of", len(rays), "brute check", brute_ok)
    np.savez_compressed(os.path.join(HERE, f"c1_{name}_golden.npz"), width=W, height=H, sub=SUB, **out)
for f in sorted(os.listdir(HERE)):
    print(f, os.path.getsize(os.path.join(HERE, f)))
