"""e2e pipeline diagnostics: rfwb200_trace_closest with pinned host buffers on C2, chunk sizes swept; with
RFWB200_PIPE_TRACE=1 the library prints the per-chunk stage completion times."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes, wire
n = 1 << 24
desc = scenes.soup_scene(1000000, 0.005)
be = backend.B200Backend(); desc.apply(be)
pr = backend.PinnedArray(n, wire.RAY); ph = backend.PinnedArray(n, wire.HIT)
pr.array[:] = scenes.random_rays(n)
for l2p, streamed in ((1, 1), (0, 1), (1, 0), (0, 0)):
    be.set_option("l2_persist", l2p)
    be.set_option("streamed", streamed)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); be.trace_closest(pr.array, out=ph.array); best = min(best, time.perf_counter() - t0)
    print(f"l2_persist={l2p} streamed={streamed}: e2e {n / best / 1e6:.1f} Mrays/s ({best * 1e3:.2f} ms), library total_ms {be.trace_stats()['total_ms']:.2f}", flush=True)
be.set_option("streamed", 0)
for chunk in [int(x) for x in os.environ.get("CHUNKS", "2097152,1048576,4194304").split(",")]:
    be.set_option("chunk_rays", chunk)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); be.trace_closest(pr.array, out=ph.array); best = min(best, time.perf_counter() - t0)
    print(f"chunk {chunk}: e2e {n / best / 1e6:.1f} Mrays/s ({best * 1e3:.2f} ms), library total_ms {be.trace_stats()['total_ms']:.2f}", flush=True)
