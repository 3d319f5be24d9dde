"""Find the first divergence between GPU and oracle for bad pixels of a render."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import backend, scenes, wire
np.set_printoptions(precision=9, linewidth=220)
sky = (0.3, 0.35, 0.5)
w, h, depth = int(os.environ.get("DBG_W", 256)), int(os.environ.get("DBG_H", 144)), 5
if os.environ.get("DBG_SCENE", "").startswith("c5:"):   # DBG_SCENE=c5:<triangles>: the C5 soup (+ ground + 64 lights) from the C5 camera
    desc = scenes.c5_scene(int(os.environ["DBG_SCENE"][3:])); view = scenes.c5_view(w, h)
else:
    view = scenes.camera_view((0, 3.5, -9.0), (0, -0.35, 1.0), w, h)
    desc = scenes.instanced_scene(grid=10, subdiv=2, n_lights=16)
orc.build()
cpu = orc.OracleBackend(det_eps=0.0); desc.apply(cpu)
gpu = backend.B200Backend(w, h, sky=sky); desc.apply(gpu)
for kv in os.environ.get("DBG_OPTS", "").split(","):   # e.g. DBG_OPTS=tri_test=1
    if "=" in kv: gpu.set_option(kv.split("=")[0], int(kv.split("=")[1]))
found = 0
for s in range(4):
    gpu.reset_accumulator(); gpu.set_option("sample_count", s); gpu.render_spp(view, 1, depth)
    a = gpu.read_accumulator()[..., :3]
    r, st = cpu.render(view, w, h, 1, depth, sky=sky, first_sample=s); r = r[..., :3]
    d = np.abs(a - r).max(axis=2)
    ys, xs = np.nonzero(d > 1e-3 * np.maximum(1.0, r.max(axis=2)))
    print(f"sample {s}: {len(ys)} of {w * h} pixels differ by more than 1e-3 (relative to max(1, radiance)); all-pixel RMSE {float(np.sqrt(((a - r) ** 2).mean())):.3e}", flush=True)
    for y, x in zip(ys, xs):
        pid = x + y * w
        probe = cpu.path_probe(view, w, h, pid, s, depth, sky=sky)
        print(f"=== sample {s} pixel ({x},{y}) gpu {a[y, x]} ref {r[y, x]}")
        for dd in range(1, depth + 1):
            gpu.reset_accumulator(); gpu.set_option("sample_count", s); gpu.render_spp(view, 1, dd)
            O, D, T, S, n = gpu.debug_read_queue(2 + ((dd - 1) & 1), w * h)
            ids = O[:n, 3].view(np.uint32)
            k = np.nonzero(ids == pid)[0]
            if len(k) == 0:
                print(f"  seg {dd - 1}: gpu path not in queue (terminated); ref row", probe[dd - 1, 6:11])
                break
            k = k[0]
            gs = S[k]
            ginst, gprim = gs[:2].view(np.int32)
            bary = gs[3:4].view(np.uint32)[0]
            print(f"  seg {dd - 1}: gpu O {O[k, :3]} D {D[k, :3]} hit ({ginst},{gprim}) t {gs[2]:.9g} u {(bary & 65535) / 65535:.6f} v {(bary >> 16) / 65535:.6f}")
            print(f"          ref O {probe[dd - 1, 0:3]} D {probe[dd - 1, 3:6]} hit ({int(probe[dd - 1, 6])},{int(probe[dd - 1, 7])}) t {probe[dd - 1, 8]:.9g} u {probe[dd - 1, 9]:.6f} v {probe[dd - 1, 10]:.6f}  acc.x after {probe[dd - 1, 22]:.6f}")
            if ginst != int(probe[dd - 1, 6]) or gprim != int(probe[dd - 1, 7]):
                rr = np.zeros(2, wire.RAY); rr["origin"][0] = O[k, :3]; rr["direction"][0] = D[k, :3]; rr["origin"][1] = probe[dd - 1, 0:3]; rr["direction"][1] = probe[dd - 1, 3:6]
                rr["tmin"] = 1e-4; rr["tmax"] = 1e26
                print("          DIVERGED. gpu ray on gpu/cpu:", gpu.trace_closest(rr[:1]), cpu.trace_closest(rr[:1]), " ref ray on gpu/cpu:", gpu.trace_closest(rr[1:]), cpu.trace_closest(rr[1:]))
                break
        found += 1
        if found >= 6:
            sys.exit(0)
