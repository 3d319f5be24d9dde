"""C3 wavefront render: sweep of the two-level traversal knobs.  env IB (inst_batch list), TB2 (tri_batch_two_level list), RF (refill_below list), SPP."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes
w, h, spp, depth = 1920, 1080, int(os.environ.get("SPP", 16)), 5
desc = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
be = backend.B200Backend(w, h, sky=(0.3, 0.35, 0.5)); desc.apply(be)
view = scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, h)
be.render_spp(view, spp, depth)
ref = None
lst = lambda k, d: [int(x) for x in os.environ.get(k, d).split(",")]
for rf in lst("RF", "28"):
    for tb in lst("TB2", "6"):
        for ib in lst("IB", "1,4,8,12,16,24"):
            be.set_option("inst_batch", ib); be.set_option("tri_batch_two_level", tb); be.set_option("refill_below", rf)
            best = 1e9
            for _ in range(3):
                be.reset_accumulator(); be.render_spp(view, spp, depth)
                best = min(best, be.render_stats()["render_ms"])
            acc = be.read_accumulator()
            if ref is None: ref = acc
            rs = be.render_stats()
            print(f"refill {rf} tri_batch_tl {tb} inst_batch {ib:2d}: {best:7.2f} ms  {rs['samples'] / best / 1e3:7.1f} Msamples/s  image identical={bool(np.array_equal(acc, ref))}", flush=True)
