"""One-off stress of the fused build kernel on the CPU tier's SIMT machine (tests/hostemu/build_emu.cpp): random sizes, treelet sizes, leaf limits and
thread counts; every primitive exactly once + hits equal to the oracle's brute force.  usage: python scripts/stress_build_emu.py FIRST LAST"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from oracle import oracle as orc
from rfw_rs_b200 import scenes
from tests import parity
from tests.test_build_emu import Built
from tests.test_hostemu import Emu, aimed_rays, degenerate_scene, load_emu

orc.build()
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "hostemu", "libbuild_emu.so"))
L.emu_build_small.restype = C.c_int
L.emu_build_small.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 6
L.emu_trace_built.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
first, last = int(sys.argv[1]), int(sys.argv[2])
bad = 0
edge_on = 0
for seed in range(first, last):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 2600), rng.integers(2600, 8193)], p=[0.2, 0.4, 0.3, 0.1]))
    threads = 512 if (n > 2048 and rng.random() < 0.7) or rng.random() < 0.1 else 256
    treelet = int(rng.choice([0, 2, 4, 8, 16, 64]))
    pmax = int(rng.integers(1, 4))
    c_prim = float(rng.choice([0.3, 0.8, 4.0]))
    dist = str(rng.choice(["soup", "identical", "two_points", "line", "plane_grid", "exponential", "huge_and_tiny"]))
    try:
        if dist == "soup":
            desc = scenes.soup_scene(n, float(rng.choice([0.3, 0.05, 0.01])), seed=seed); tris = desc.meshes[0]
        else:
            desc, tris, _ = degenerate_scene(dist, n)
        b = Built(L, tris, treelet=treelet, threads=threads, c_prim=c_prim, pmax=pmax)
        assert b.rc == 0, f"rc {b.rc}"
        assert int(b.counters[1]) == len(tris) and sorted(b.leaf.tolist()) == list(range(len(tris)))
        o = orc.OracleBackend(det_eps=0.0); desc.apply(o)
        rays = aimed_rays(tris, np.random.default_rng(seed + 1), len(tris), count=400)
        hb, ref = b.trace(rays), o.trace_closest(rays, mode=orc.MODE_BRUTE)
        try:
            parity.compare_hits(rays, hb, ref, parity.lookup_from_desc(desc), f"seed {seed}", max_fraction=2e-2, oracle_artefacts=True)
        except AssertionError:
            # not the builder's doing if the SERIAL harness build of the same mesh (another tree: no refinement) returns the very same hits: then the
            # disagreement is between the watertight test and the oracle's float32 Moller-Trumbore on a triangle seen edge-on (tiny triangles of the
            # exponential / huge_and_tiny distributions, aimed at: projected area ~1e-11, the coordinate rounding decides the sign of an edge function)
            hs, _, _ = Emu(load_emu(), desc).trace(rays)
            if not np.array_equal(hs["prim"], hb["prim"]):
                raise
            edge_on += 1
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("seed", seed, (n, threads, treelet, pmax, c_prim, dist), "FAILED:", repr(e)[:300], flush=True)
print(f"fused build kernel on the SIMT machine, seeds {first}..{last - 1}: {bad} failures ({edge_on} seeds with edge-on triangle-test disagreements the serial build reproduces hit for hit)", flush=True)
