"""CPU tier: the product's FUSED BUILD KERNEL itself (rfw_rs_b200/csrc/build_small.cuh::k_build_small — one CTA builds one mesh: boxes, bounds, Morton,
the in-CTA radix sort, Karras, the round-based fit + cost DP, the binned-SAH top build, the collapse, the traversal triangles) compiled for the host and run
on the lane-thread SIMT machine (tests/hostemu/build_emu.cpp, simt_machine.h): one std::thread per CUDA thread, warp collectives — full and masked —
__syncthreads and the atomics are real synchronisations, so a missing barrier or a race of the CTA-scope schedule hangs or corrupts HERE, without a GPU.
Checked: every triangle lands in exactly one leaf slot, the tree's hits equal the oracle's, and with the refinement off the tree costs exactly what the
serial harness build of the same bodies (tests/hostemu/emu.cpp) costs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from rfw_rs_b200 import scenes, wire
from tests import parity
from tests.test_hostemu import Emu, aimed_rays, degenerate_scene, load_emu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def bemu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hostemu")])
    L = C.CDLL(os.path.join(HERE, "hostemu", "libbuild_emu.so"))
    L.emu_build_small.restype = C.c_int
    L.emu_build_small.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 6
    L.emu_trace_built.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    return L


class Built:
    def __init__(self, L, tris, treelet, threads, c_prim=0.8, pmax=3):
        self.L, self.n = L, len(tris)
        self.tris = np.ascontiguousarray(tris)
        self.nodes = np.zeros((self.n * 6, 4), np.float32)   # NODE_F4 = 6 float4 per wide node (96-byte stride)
        self.leaf = np.full(self.n, 0xFFFFFFFF, np.uint32)
        self.ttris = np.zeros((3 * self.n, 4), np.float32)
        self.counters = np.zeros(8, np.uint32); self.bounds = np.zeros(12, np.uint32); self.cost = np.zeros(8, np.float32)
        self.rc = L.emu_build_small(self.tris.ctypes.data, self.n, treelet, c_prim, pmax, threads, self.nodes.ctypes.data, self.leaf.ctypes.data, self.ttris.ctypes.data,
                                    self.counters.ctypes.data, self.bounds.ctypes.data, self.cost.ctypes.data)

    def trace(self, rays):
        hits = np.zeros(len(rays), wire.HIT)
        self.L.emu_trace_built(self.nodes.ctypes.data, self.ttris.ctypes.data, rays.ctypes.data, len(rays), hits.ctypes.data)
        return hits


# (n, threads): the special cases of the pipeline — no Karras tree, no refinement, one warp, one sort tile and one past it (256 threads: the sort goes
# multi-tile at 2 049), a medium job on 512 threads (two tiles of 4 096)
CASES = [(1, 256), (2, 256), (3, 256), (8, 256), (9, 256), (33, 256), (257, 256), (1000, 256), (2048, 256), (2049, 256), (777, 512), (4500, 512)]


@pytest.mark.parametrize("n,threads", CASES)
def test_fused_build_kernel_on_the_simt_machine(bemu, oracle_mod, n, threads):
    s = 0.3 if n < 200 else 0.06
    desc = scenes.soup_scene(n, s, seed=scenes.SEED_SCENE + n)
    tris = desc.meshes[0]
    b = Built(bemu, tris, treelet=8, threads=threads)
    assert b.rc == 0, "a thread missed a barrier / warp collective (hang)" if b.rc == -1 else b.rc
    # every triangle in exactly one leaf slot; the traversal triangles are the leaf order of the mesh
    assert int(b.counters[1]) == n and sorted(b.leaf.tolist()) == list(range(n))
    assert 1 <= int(b.counters[0]) <= max(1, n)
    assert np.array_equal(b.ttris[0::3, 3].view(np.uint32), b.leaf)
    assert np.array_equal(b.ttris[0::3, :3], tris["vertex0"][b.leaf]) and np.array_equal(b.ttris[2::3, :3], tris["vertex2"][b.leaf])
    if n > 8:
        assert int(b.counters[4]) >= 2   # the binned-SAH top build ran over that many treelets
    # hits through the built tree (the product's per-ray loop) against the oracle
    cpu = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(cpu)
    rays = scenes.random_rays(3000 if n > 1000 else 1500, seed=scenes.SEED_RAYS + n)
    ref = cpu.trace_closest(rays, mode=oracle_mod.MODE_BVH2)
    parity.compare_hits(rays, b.trace(rays), ref, parity.lookup_from_desc(desc), f"fused build kernel on the CPU, n = {n}")
    if n >= 100:
        assert (ref["inst"] >= 0).mean() > 0.05


@pytest.mark.parametrize("n,threads", [(2, 256), (40, 256), (700, 256), (2500, 256), (3000, 512)])
def test_fused_build_without_refinement_costs_what_the_serial_build_costs(bemu, n, threads):
    """treelet 0 = plain LBVH + SAH collapse: the serial harness (emu.cpp: the same bodies, one after the other) and the kernel under the SIMT machine
    build the same tree — same SAH cost to the bit, same number of wide nodes."""
    desc = scenes.soup_scene(n, 0.1, seed=scenes.SEED_SCENE + 3 * n)
    serial = Emu(load_emu(), desc)    # c_prim 0.3, pmax 3, no refinement
    b = Built(bemu, desc.meshes[0], treelet=0, threads=threads, c_prim=0.3, pmax=3)
    assert b.rc == 0
    sah = float(b.cost[0] / b.cost[7]) if b.cost[7] > 0 else 0.0
    assert np.float32(sah) == np.float32(serial.L.emu_sah(serial.h, 0))
    assert int(b.counters[0]) == int(serial.stats[0])


def test_fused_build_is_deterministic_under_thread_scheduling(bemu):
    """Node slots are handed out by atomics (layout differs from run to run), the TREE does not: two runs of the kernel — different thread interleavings on the
    host — give the same multiset of node records once the two allocation-order dependent words are left out, and the same leaf order per node."""
    desc = scenes.soup_scene(600, 0.1)
    a = Built(bemu, desc.meshes[0], treelet=8, threads=256)
    b = Built(bemu, desc.meshes[0], treelet=8, threads=256)
    assert a.rc == 0 and b.rc == 0 and int(a.counters[0]) == int(b.counters[0])
    k = int(a.counters[0])

    def records(x):
        w = x.nodes.reshape(-1, 24)[:k].view(np.uint32).copy()
        w[:, 4] = 0; w[:, 5] = 0   # child_base / prim_base: allocation order
        return sorted(map(bytes, w))

    assert records(a) == records(b)
    assert np.float32(a.cost[0]) == np.float32(b.cost[0])


@pytest.mark.parametrize("dist", ["identical", "two_points", "line", "plane_grid", "exponential", "huge_and_tiny"])
@pytest.mark.parametrize("n,threads", [(9, 256), (257, 256), (2300, 256), (5000, 512)])
def test_fused_build_kernel_on_degenerate_distributions(bemu, oracle_mod, dist, n, threads):
    """The inputs that stress the Morton / sort / Karras / collapse logic (all Morton keys equal: every radix pass puts all keys into ONE digit and the tree
    degenerates to index splits; two clusters; a line; a planar grid; exponentially spread sizes; one huge triangle among tiny ones) through the fused kernel
    under the SIMT machine: every primitive once, hits equal to the brute force."""
    desc, tris, rng = degenerate_scene(dist, n)
    b = Built(bemu, tris, treelet=8, threads=threads)
    assert b.rc == 0
    assert int(b.counters[1]) == len(tris) and sorted(b.leaf.tolist()) == list(range(len(tris)))
    o = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(o)
    rays = aimed_rays(tris, rng, len(tris), count=800)
    ref = o.trace_closest(rays, mode=oracle_mod.MODE_BRUTE)
    parity.compare_hits(rays, b.trace(rays), ref, parity.lookup_from_desc(desc), f"fused {dist}/{n}", max_fraction=2e-2, oracle_artefacts=True)


@pytest.mark.parametrize("n,threads,bits", [(0, 256, (0, 64)), (1, 256, (0, 64)), (31, 256, (0, 64)), (2048, 256, (0, 64)), (2049, 256, (0, 64)), (5000, 256, (0, 64)),
                                            (8192, 256, (0, 24)), (3000, 512, (0, 64)), (4097, 512, (0, 64)), (8192, 512, (8, 40)), (700, 256, (0, 8))])
def test_in_cta_radix_sort_on_the_simt_machine(bemu, n, threads, bits):
    """sort_small.cuh::sort_tiles_body alone: one tile (keys stay in registers) and several tiles (two sweeps per pass, a base per digit and tile), 8 and 16 warps,
    full and partial bit ranges (odd pass counts end in the tmp buffers), many duplicate keys — against numpy's stable sort on the same bit range."""
    bemu.emu_sort_pairs.restype = C.c_int
    bemu.emu_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(n + threads)
    keys = rng.integers(0, 1 << 63, max(n, 1), dtype=np.uint64)[:n].copy()
    if n > 10:
        keys[rng.integers(0, n, n // 3)] = keys[0]                     # a third of the keys equal
        keys[rng.integers(0, n, n // 5)] &= np.uint64(0xFFFF)          # and a cluster in the low digits
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    assert bemu.emu_sort_pairs(k.ctypes.data, v.ctypes.data, n, threads, bits[0], bits[1]) == 0
    mask = np.uint64(((1 << (bits[1] - bits[0])) - 1) << bits[0]) if bits[1] - bits[0] < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(v, vals[order]) and np.array_equal(k, keys[order])


@pytest.mark.parametrize("n", [300, 1500, 2600])
def test_cta_cooperative_sah_split_builds_the_same_tree(bemu, n):
    """builder: middle-sized segments of the SAH top build are split by ALL warps of a CTA in the cooperative k_sah_top (sah_split_segment_cta) — claimed to give
    the tree the one-warp split gives.  The fused kernel instantiated with a scope that sends every segment of more than 32 treelets through that function, against
    the shipped instantiation (one warp per segment): same SAH cost to the bit, same multiset of node records."""
    desc = scenes.soup_scene(n, 0.08, seed=scenes.SEED_SCENE + 7 * n)
    a = Built(bemu, desc.meshes[0], treelet=4, threads=256)
    b = Built(bemu, desc.meshes[0], treelet=4, threads=-256)
    assert a.rc == 0 and b.rc == 0
    assert int(a.counters[4]) > 64                      # enough treelets for several cooperative levels
    assert a.counters.tolist()[:2] == b.counters.tolist()[:2] and int(a.counters[4]) == int(b.counters[4])
    assert np.array_equal(a.cost.view(np.uint32), b.cost.view(np.uint32))
    k = int(a.counters[0])

    def records(x):
        w = x.nodes.reshape(-1, 24)[:k].view(np.uint32).copy()
        w[:, 4] = 0; w[:, 5] = 0
        return sorted(bytes(r) for r in w)

    assert records(a) == records(b)
