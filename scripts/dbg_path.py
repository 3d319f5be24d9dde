import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes, wire
np.set_printoptions(precision=8, suppress=False, linewidth=200)
desc = scenes.instanced_scene(grid=6, subdiv=1, n_lights=4)
w, h = 96, 54
sky = (0.2, 0.2, 0.3)
view = scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h)
cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
px, py, sample = 36, 26, 1
pid = px + py * w
probe = cpu.path_probe(view, w, h, pid, sample, 3, sky=sky)
print("oracle probe rows:\n", probe)
gpu = backend.B200Backend(w, h, sky=sky); desc.apply(gpu)
for depth in (1, 2):
    gpu.reset_accumulator(); gpu.set_option("sample_count", sample)
    gpu.render_spp(view, 1, depth)
    which = depth & 1
    O, D, T, S, n = gpu.debug_read_queue(which, w * h)
    ids = O[:n, 3].view(np.uint32)
    k = np.nonzero(ids == pid)[0]
    print("depth", depth, "queue", which, "count", n, "found", k)
    for i in k:
        print("  gpu next O", O[i, :3], "D", D[i, :3], "T", T[i])
        print("  ref next O", probe[depth - 1, 16:19], "D", probe[depth - 1, 19:22], "T/pdf", probe[depth, 11:15] if depth < 3 else None)
# trace the oracle's 3rd-segment ray on both
r = np.zeros(1, wire.RAY); r["origin"] = probe[2, 0:3]; r["direction"] = probe[2, 3:6]; r["tmin"] = 1e-4; r["tmax"] = 1e26
print("3rd segment ray: gpu", gpu.trace_closest(r), "ref", cpu.trace_closest(r), "probe", probe[2, 6:11])
