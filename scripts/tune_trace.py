"""Sweep the persistent-kernel tuning knobs on config C2 (1M-triangle soup, 2^24 incoherent rays)."""
import sys, os, json, itertools; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfw_rs_b200 import backend, scenes, wire
n_tris = int(os.environ.get("TUNE_TRIS", 1000000)); s = float(os.environ.get("TUNE_S", 0.005)); n_rays = int(os.environ.get("TUNE_RAYS", 1 << 24))
desc = scenes.soup_scene(n_tris, s)
be = backend.B200Backend(); desc.apply(be)
print("build", be.build_stats())
if "SORT_RAYS" in os.environ: be.set_option("sort_rays", int(os.environ["SORT_RAYS"]))
rays = scenes.random_rays(n_rays)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
d_occ = torch.empty(n_rays, dtype=torch.int32, device="cuda")
def run(any_hit=False, reps=3):
    best = 1e9
    for _ in range(reps):
        if any_hit: be.trace_any_device(d_rays.data_ptr(), n_rays, d_occ.data_ptr())
        else: be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr())
        best = min(best, be.trace_stats()["kernel_ms"])
    return n_rays / best / 1e3
be.set_option("trace_variant", 1); print("simple kernel closest Mrays/s", run(), "any", run(True)); be.set_option("trace_variant", 0)
ref = None
results = []
mbs = [int(x) for x in os.environ.get("TUNE_MB", "4,5,6,8").split(",")]
tbs = [int(x) for x in os.environ.get("TUNE_TB", "1,8,12,16,20,24").split(",")]
rfs = [int(x) for x in os.environ.get("TUNE_RF", "16,24,28").split(",")]
bls = [int(x) for x in os.environ.get("TUNE_BL", "4").split(",")]
for mb, tb, rf, bl in itertools.product(mbs, tbs, rfs, bls):
    be.set_option("min_blocks", mb); be.set_option("tri_batch", tb); be.set_option("refill_below", rf); be.set_option("tri_blocked", bl)
    r = run(reps=2)
    h = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT)
    if ref is None: ref = h.copy()
    same = np.array_equal(ref["prim"], h["prim"]) and np.array_equal(ref["t"], h["t"])
    results.append((r, mb, tb, rf, same, bl))
    print(f"min_blocks {mb} tri_batch {tb:2d} tri_blocked {bl:2d} refill_below {rf:2d}: closest {r:8.1f} Mrays/s  same={same}", flush=True)
results.sort(reverse=True)
print("BEST", results[:5])
r, mb, tb, rf, _, bl = results[0]
be.set_option("min_blocks", mb); be.set_option("tri_batch", tb); be.set_option("refill_below", rf); be.set_option("tri_blocked", bl)
print("any-hit at best closest config:", run(True))
