"""What rfw's render loop sees: Backend::render(view, Default) = 1 spp per call at window resolution (C3 scene, 1280x720 and 1920x1080),
wall clock per frame incl. finalize, with a camera that moves every frame (accumulation restarts) and one that stands still."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes
desc = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
for (w, h) in ((1280, 720), (1920, 1080)):
    be = backend.B200Backend(w, h, sky=(0.3, 0.35, 0.5)); desc.apply(be)
    views = [scenes.camera_view((0.02 * k, 14.0, -62.0), (0.0, -0.25, 1.0), w, h) for k in range(60)]
    for v in views[:5]: be.render(None, v, 0)
    l0 = be.launch_count(); t0 = time.perf_counter()
    for v in views[5:55]: be.render(None, v, 0)
    moving = (time.perf_counter() - t0) / 50 * 1e3; per_frame = (be.launch_count() - l0) / 50
    t0 = time.perf_counter()
    for _ in range(50): be.render(None, views[0], 0)
    still = (time.perf_counter() - t0) / 50 * 1e3
    print(f"{w}x{h}: moving camera {moving:.2f} ms/frame ({1e3 / moving:.0f} fps), still camera {still:.2f} ms/frame, {per_frame:.0f} kernel launches per frame, depth {3}", flush=True)
    # moving instances: TLAS rebuild + render per frame
    import copy
    mats = {m: np.array(desc.instances[m], np.float32).copy() for m in range(8)}
    t0 = time.perf_counter()
    for k in range(30):
        for m in range(8):
            a = mats[m].reshape(-1, 16).copy(); a[:, 13] += 0.01 * np.sin(0.3 * k + m); be.set_3d_instances(m, a)
        be.synchronize(); be.render(None, views[0], 0)
    dyn = (time.perf_counter() - t0) / 30 * 1e3
    print(f"{w}x{h}: all 10 000 sphere instances moved every frame (8 x set_3d_instances + synchronize + render): {dyn:.2f} ms/frame ({1e3 / dyn:.0f} fps)", flush=True)
