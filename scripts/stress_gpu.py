"""The randomised builder + traversal stress of tests/test_hostemu.py (awkward scales, flat / duplicate / sliver triangles, scaled instances,
axis-parallel and on-surface rays, against the oracle's brute force over all triangles) with the GPU library in place of the host harness.
usage: python scripts/stress_gpu.py FIRST_SEED LAST_SEED"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from rfw_rs_b200 import backend
from tests import test_hostemu as T


class GpuEmu:
    def __init__(self, _lib, desc):
        self.be = backend.B200Backend(); desc.apply(self.be)

    def trace(self, rays):
        return self.be.trace_closest(rays), self.be.trace_any(rays), None


def run(first, last):
    orc.build()
    T.Emu = GpuEmu
    fn = T.test_builder_and_traversal_stress_against_brute_force
    bad = 0
    for seed in range(first, last):
        try:
            fn(None, orc, seed)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("seed", seed, "FAILED:", repr(e)[:300], flush=True)
    print(f"GPU stress seeds {first}..{last - 1}: {bad} failures", flush=True)
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]), int(sys.argv[2])) else 0)
