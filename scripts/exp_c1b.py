"""C1b (pica, 170 BLAS + TLAS, 1280x720 primary rays): which knob moved it?"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, gltf
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
desc = gltf.per_mesh(asset); w, h = 1280, 720
view = gltf.c1_camera(gltf.flatten(asset), w, h)
def run(label, **opts):
    be = backend.B200Backend(w, h)
    for k, v in opts.items(): be.set_option(k, v)
    desc.apply(be)
    best = 1e9
    for _ in range(8):
        be.cast_primary(view); best = min(best, be.trace_stats()["kernel_ms"])
    print(f"{label:40s} {best*1e3:7.1f} us  {w*h/best/1e3:8.0f} Mrays/s  nodes {be.build_stats()['blas_nodes']}", flush=True)
run("defaults")
run("treelet 0 (both)", sah_treelet=0, sah_treelet_tlas=0)
run("BLAS refined, TLAS plain LBVH", sah_treelet=8, sah_treelet_tlas=0)
run("BLAS plain, TLAS refined", sah_treelet=0, sah_treelet_tlas=8)
run("BLAS treelet 32, TLAS 0", sah_treelet=32, sah_treelet_tlas=0)
run("BLAS treelet 4, TLAS 4", sah_treelet=4, sah_treelet_tlas=4)
