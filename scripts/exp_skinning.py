"""Per-frame cost of animated (skinned) characters: COPIES instances of CesiumMan (4 672 triangles) share one mesh, all but the last are skinned;
every frame uploads a new pose (set_skins) and synchronize() re-skins + rebuilds one BLAS per skinned instance + the TLAS."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, gltf, scenes
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
a = gltf.load_npz(os.path.join(root, "tests", "golden", "cesium_man.npz"))
for copies in [int(x) for x in os.environ.get("COPIES", "2,17,65").split(",")]:
    sc = gltf.skinned(a, copies=copies)
    be = backend.B200Backend()
    for kv in os.environ.get("OPTS", "").split(","):
        if "=" in kv: be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    sc.apply(be)
    rays = scenes.random_rays(1 << 16, lo=-1.0, hi=1.0)
    times = []
    for f in range(12):
        pose = gltf.pose_joints(a.skins[0], angle=0.1 + 0.03 * f)
        l0 = be.launch_count(); t0 = time.perf_counter()
        be.set_skins([pose]); be.synchronize()
        times.append((time.perf_counter() - t0) * 1e3); nl = be.launch_count() - l0
    h = be.trace_closest(rays)
    import zlib
    print(f"copies {copies:3d} (skinned {copies - 1:3d}): frame set_skins + synchronize wall ms min {min(times[2:]):7.3f} median {sorted(times[2:])[len(times[2:]) // 2]:7.3f}; "
          f"launches {nl}; hit crc {zlib.crc32(h.tobytes()):08x} hit rate {float((h['inst'] >= 0).mean()):.3f}", flush=True)
