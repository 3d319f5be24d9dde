"""Where do hits with spatial splits differ from hits without (pica, flattened)?"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes, wire, gltf
from oracle import oracle as orc
from tests import parity
orc.build()
asset = gltf.load_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pica.npz"))
flat = gltf.flatten(asset)
tris = flat.meshes[0]
lo = tris["vertex0"].min(axis=0); hi = tris["vertex0"].max(axis=0)
n = 1 << 20
rays = scenes.random_rays(n, lo=0.0, hi=1.0); rays["origin"] = lo + rays["origin"] * (hi - lo)
res = {}
for b in (0, 30):
    be = backend.B200Backend(); be.set_option("split_budget", b); flat.apply(be)
    res[b] = be.trace_closest(rays)
a, c = res[0], res[30]
diff = np.nonzero((a["prim"] != c["prim"]) | (a["t"] != c["t"]))[0]
print("rays", n, "differ", len(diff), "of which prim differs", int((a["prim"][diff] != c["prim"][diff]).sum()))
miss_only_split = diff[(c["prim"][diff] < 0) & (a["prim"][diff] >= 0)]
miss_only_plain = diff[(a["prim"][diff] < 0) & (c["prim"][diff] >= 0)]
print("hit without splits but MISS with:", len(miss_only_split), " hit with splits but miss without:", len(miss_only_plain))
both = diff[(a["prim"][diff] >= 0) & (c["prim"][diff] >= 0)]
if len(both):
    rel = np.abs(a["t"][both] - c["t"][both]) / np.maximum(a["t"][both], 1e-30)
    print("both hit, different: ", len(both), "relative t difference: median", float(np.median(rel)), "max", float(rel.max()), " split closer:", int((c["t"][both] < a["t"][both]).sum()), " plain closer:", int((a["t"][both] < c["t"][both]).sum()))
for i in diff[:8]:
    print(i, "plain", a[i], "split", c[i])
o = orc.OracleBackend(det_eps=0.0); flat.apply(o)
ref = o.trace_closest(rays[:1 << 17], mode=orc.MODE_BVH2)
for b in (0, 30):
    try:
        nb = parity.compare_hits(rays[:1 << 17], res[b][:1 << 17], ref, parity.lookup_from_desc(flat), f"pica budget {b}", max_fraction=1e-2)
        print("budget", b, "classified near-ties vs oracle:", nb)
    except AssertionError as e:
        print("budget", b, "PARITY FAILURE:", str(e)[:400])
