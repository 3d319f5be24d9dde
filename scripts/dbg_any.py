import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes, wire
desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
gpu = backend.B200Backend(); desc.apply(gpu)
# NEE-like rays: from points just above the ground to points on the lights (y = 6)
n = 200000
r = scenes.u01(7, np.arange(n * 6).reshape(n, 6)).astype(np.float64)
o = np.stack([r[:, 0] * 7 - 3.5, 1.5e-5 + 0 * r[:, 1], r[:, 2] * 7 - 3.5], axis=1)
p = np.stack([r[:, 3] * 7 - 3.5, np.full(n, 6.0), r[:, 5] * 7 - 3.5], axis=1)
d = p - o; dist = np.linalg.norm(d, axis=1); d /= dist[:, None]
rays = np.zeros(n, wire.RAY); rays["origin"] = o; rays["direction"] = d; rays["tmin"] = 1e-3; rays["tmax"] = dist - 2e-4
for variant in (0, 1):
    gpu.set_option("trace_variant", variant)
    g = gpu.trace_any(rays); c = cpu.trace_any(rays)
    bad = np.nonzero((g != 0) != (c != 0))[0]
    print("variant", variant, "occluded frac gpu/cpu", (g != 0).mean(), (c != 0).mean(), "mismatch", len(bad))
    hc = cpu.trace_closest(rays); hg = gpu.trace_closest(rays)
    print("   closest mismatches", ((hc["inst"] != hg["inst"]) | (hc["prim"] != hg["prim"])).sum())
    for i in bad[:5]:
        print("   ray", i, rays[i], "gpu any", g[i], "cpu any", c[i], "cpu closest", hc[i], "gpu closest", hg[i])
