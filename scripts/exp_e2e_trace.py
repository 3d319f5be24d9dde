"""One streamed e2e pass on C2 with the library's pipeline trace (RFWB200_PIPE_TRACE=1) + pure-copy ceilings."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n = 1 << 24
desc = scenes.soup_scene(1000000, 0.005)
be = backend.B200Backend(); desc.apply(be)
pr = backend.PinnedArray(n, wire.RAY); ph = backend.PinnedArray(n, wire.HIT)
pr.array[:] = scenes.random_rays(n)
for k in range(6):
    t0 = time.perf_counter(); be.trace_closest(pr.array, out=ph.array); dt = time.perf_counter() - t0
    print(f"pass {k}: e2e {n / dt / 1e6:.1f} Mrays/s ({dt * 1e3:.2f} ms wall), library total_ms {be.trace_stats()['total_ms']:.2f}", flush=True)
# pure copies of the same buffers (torch streams): H2D alone, D2H alone, both at once
hr = torch.from_numpy(pr.array.view(np.uint8).reshape(-1)); hh = torch.from_numpy(ph.array.view(np.uint8).reshape(-1))
dr = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); dh = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
def h2d():
    with torch.cuda.stream(s1): dr.copy_(hr, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hh.copy_(dh, non_blocking=True)
def both(): h2d(); d2h()
print(f"H2D 512 MiB alone {timed(h2d):.2f} ms, D2H 320 MiB alone {timed(d2h):.2f} ms, both at once {timed(both):.2f} ms")
