"""Warm rebuild of pica's 170 per-mesh BLASes (asset scenes submit one Mesh3D per glTF mesh)."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, gltf
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
desc = gltf.per_mesh(asset)
be = backend.B200Backend(); be.set_option("build_fused", int(os.environ.get("BUILD_FUSED", 1)))
for kv in os.environ.get("OPTS", "").split(","):
    if "=" in kv: be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
desc.apply(be)
print("cold", be.build_stats()["blas_build_ms"], "ms; meshes", be.build_stats()["num_meshes"], "launches", be.launch_count())
for k in range(8):
    be.set_option("build_streams", 1 if k < 4 else int(os.environ.get("BUILD_STREAMS", 8)))  # first four: everything on the main stream
    be.set_option("build_threads", int(os.environ.get("BUILD_THREADS", 1)))
    l0 = be.launch_count()
    be.set_option("sah_treelet", 8); t0 = time.perf_counter(); be.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"warm rebuild {k}: blas_build_ms {be.build_stats()['blas_build_ms']:.2f} (device events), synchronize wall {dt:.2f} ms, kernel launches {be.launch_count() - l0}")
sizes = sorted(len(t) for t in desc.meshes.values())
from rfw_rs_b200 import scenes
rays = scenes.random_rays(1 << 18, lo=-1.0, hi=1.0)
h8 = be.trace_closest(rays)
be.set_option("build_streams", 1); be.set_option("sah_treelet", 8); be.synchronize()
h1 = be.trace_closest(rays)
print("hits identical across build_streams:", bool(np.array_equal(h1, h8)), "hit rate", float((h1["inst"] >= 0).mean()))
print("mesh sizes: min", sizes[0], "median", sizes[len(sizes)//2], "max", sizes[-1])
