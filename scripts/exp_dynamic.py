"""Dynamic scene: per-frame instance-matrix updates -> device instance records + TLAS-only rebuild.
Pattern of examples/animated/src/main.rs:197-219 (every instance moves every frame)."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes
grid = int(os.environ.get("GRID", 100))
desc = scenes.instanced_scene(grid=grid, subdiv=3, n_lights=16)
w, h = 1280, 720
be = backend.B200Backend(w, h, sky=(0.3, 0.35, 0.5))
for kv in os.environ.get("OPTS", "").split(","):
    if "=" in kv: be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
desc.apply(be)
view = scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, h)

times, tl = [], []
for frame in range(12):
    for m in range(8):
        M = desc.instances[m].reshape(-1, 4, 4).copy()           # column-major: translation in [:, 3, :3]
        M[:, 3, 1] = 0.4 + 0.3 * np.abs(np.sin(0.3 * frame + np.arange(len(M))))
        t0 = time.perf_counter(); be.set_3d_instances(m, M.reshape(-1, 16)); times.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); be.synchronize(); dt = time.perf_counter() - t0
    bs = be.build_stats()
    tl.append((dt * 1e3, bs["tlas_build_ms"], bs["blas_build_ms"]))
    be.render(None, view)
print("per-frame synchronize wall ms / tlas device ms / blas ms:", [(round(a, 3), round(b, 3), round(c, 3)) for a, b, c in tl])
print(f"median synchronize wall {np.median([a for a, _, _ in tl[2:]]):.3f} ms for {bs['num_instances']} instances; set_3d_instances total per frame {np.sum(times) / 12 * 1e3:.3f} ms")
