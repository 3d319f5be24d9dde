"""Sparse scene (most rays miss): host-streamed single launch vs chunked pipeline, single- and two-level.  A lane parked with
pending triangles / a pending instance entry must not hold back the in-order download watermark while the other lanes of
its warp churn through missing rays."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes, wire
n = 1 << 23
def scene(two_level):
    sc = scenes.SceneDesc(); sc.materials = scenes.material()
    t = scenes.soup(4000, 0.004)
    t2 = t.copy()
    for k in ("vertex0", "vertex1", "vertex2"): t2[k] = t[k] * 0.08 + 0.46   # a small cluster in the middle of the unit cube
    sc.meshes[0] = scenes.make_triangles(t2["vertex0"], t2["vertex1"], t2["vertex2"])
    sc.instances[0] = scenes.to_column_major([scenes.identity()] if not two_level else [scenes.identity(), scenes.trs((0.3, 0.3, 0.3)), scenes.trs((-0.3, 0.2, -0.3))])
    return sc
for two_level in (False, True):
    be = backend.B200Backend(); scene(two_level).apply(be)
    pr = backend.PinnedArray(n, wire.RAY); ph = backend.PinnedArray(n, wire.HIT); pr.array[:] = scenes.random_rays(n)
    res = {}
    for streamed in (1, 0):
        be.set_option("streamed", streamed)
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter(); be.trace_closest(pr.array, out=ph.array); best = min(best, time.perf_counter() - t0)
        res[streamed] = ph.array.copy()
        print(f"two_level={two_level} streamed={streamed}: {n / best / 1e6:.0f} Mrays/s ({best * 1e3:.2f} ms), hit rate {(ph.array['inst'] >= 0).mean():.4f}", flush=True)
    print("identical:", bool(np.array_equal(res[0], res[1])))
