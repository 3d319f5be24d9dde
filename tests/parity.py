"""Parity policy shared by the CPU-tier logic tests and the -m gpu tests.

IDs must be bit-exact except at documented near-ties (BASELINE.json north_star).  The product uses a
watertight (Woop) triangle test while the oracle restates the reference's Möller–Trumbore
(intersection.glsl:1-38); the two can disagree only when a ray passes within float rounding of a triangle
edge, or when two hits are closer in t than the formulations' rounding.  `compare_hits` checks:
  * t within REL_T (1e-4 relative) and barycentrics within ABS_UV wherever the IDs agree,
  * every ID disagreement is classified, in float64, as a near-tie (|t_a - t_b| <= REL_T * t) or an
    edge graze (the winning triangle of one side is hit within EDGE_EPS of one of its edges),
  * the number of such disagreements is bounded by MAX_MISMATCH_FRACTION of the rays.
"""
import numpy as np

REL_T = 1e-4
ULPS_T = 16  # absolute floor: float32 cannot resolve t finer than a few ulp of the hit point's coordinates
ABS_UV = 2e-3
EDGE_EPS = 1e-4
MAX_MISMATCH_FRACTION = 2e-5


def _tri_f64(tris, prim):
    return (tris["vertex0"][prim].astype(np.float64), tris["vertex1"][prim].astype(np.float64), tris["vertex2"][prim].astype(np.float64))


def exact_hit(o, d, v0, v1, v2):
    """float64 Möller–Trumbore: returns (t, u, v) or None when parallel."""
    e1, e2 = v1 - v0, v2 - v0
    h = np.cross(d, e2)
    a = np.dot(e1, h)
    if a == 0:
        return None
    f = 1.0 / a
    s = o - v0
    u = f * np.dot(s, h)
    q = np.cross(s, e1)
    v = f * np.dot(d, q)
    t = f * np.dot(e2, q)
    return t, u, v


def t_tolerance(rays, t):
    """|dt| <= 1e-4 * t  +  ULPS_T ulp(float32) of the largest coordinate along the segment, in units of t."""
    o = rays["origin"].astype(np.float64)
    d = rays["direction"].astype(np.float64)
    t = np.asarray(t, np.float64)
    scale = np.maximum(np.abs(o).max(axis=1), np.abs(o + d * t[:, None]).max(axis=1))
    dlen = np.maximum(np.linalg.norm(d, axis=1), 1e-30)
    return REL_T * np.abs(t) + ULPS_T * 2.0 ** -24 * scale / dlen


def mt_determinant(o, d, v0, v1, v2):
    """float64 value of the determinant the reference compares with its epsilon (intersection.glsl:10-12): dot(edge1, cross(d, edge2))"""
    return float(np.dot(v1 - v0, np.cross(d, v2 - v0)))


def edge_on(o, d, v0, v1, v2, ulps=64.0):
    """Is the triangle's projection along the ray thinner than `ulps` float32 ulps of its coordinates relative to the ray origin?"""
    dn = d / max(np.linalg.norm(d), 1e-300)
    P = [(v - o) - np.dot(v - o, dn) * dn for v in (v0, v1, v2)]   # the vertices projected onto the plane perpendicular to the ray
    area2 = np.linalg.norm(np.cross(P[1] - P[0], P[2] - P[0]))
    longest = max(np.linalg.norm(P[1] - P[0]), np.linalg.norm(P[2] - P[1]), np.linalg.norm(P[0] - P[2]), 1e-300)
    scale = max(np.abs(v0 - o).max(), np.abs(v1 - o).max(), np.abs(v2 - o).max())
    return area2 / longest <= ulps * 2.0 ** -24 * scale


def compare_hits(rays, gpu, ref, scene_lookup, label="", max_fraction=MAX_MISMATCH_FRACTION, oracle_artefacts=False, reference_epsilon=0.0):
    """max_fraction bounds the number of CLASSIFIED near-ties (unclassified ones are never allowed).  Authored assets
    with coplanar duplicated faces (pica) need a looser count bound than the synthetic scenes: every pixel looking at
    such a face pair is an exact-depth tie between two different triangles."""
    """scene_lookup(inst) -> (tris, 4x4 inverse matrix as float64 row-indexed) for the global instance id.
    reference_epsilon > 0 (comparisons with hits produced by the reference's own intersection.glsl, which rejects triangles whose
    Möller-Trumbore determinant is below 1e-4 in magnitude — small or grazing triangles): a product hit that is exact-valid, not
    farther than the reference's, on a triangle whose float64 determinant is below that epsilon is explained by it.
    oracle_artefacts=True (stress tests on ill-conditioned inputs only): a hit the ORACLE reports whose float64-exact t lies outside
    the ray's (tmin, tmax) — its float32 Moller-Trumbore on a sliver triangle with the origin on the surface — explains a
    mismatch as well; the product's answer is then checked to be exact-valid (or a miss); so does a triangle seen edge-on (edge_on())."""
    n = len(rays)
    assert len(gpu) == n and len(ref) == n
    same = (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"])
    hit = same & (ref["inst"] >= 0)
    if hit.any():
        idx = np.nonzero(hit)[0]
        tol = t_tolerance(rays[idx], ref["t"][idx])
        err = np.abs(gpu["t"][idx].astype(np.float64) - ref["t"][idx].astype(np.float64))
        # grazing hits: a float32 perturbation of the ray moves t by 1/|cos(theta)| as much as it moves the hit point
        # off the surface, so for the (few) hits over the plain tolerance the error is measured along the normal
        for j in np.nonzero(err > tol)[0]:
            i = idx[j]
            tris, inv = scene_lookup(int(ref["inst"][i]))
            v0, v1, v2 = _tri_f64(tris, int(ref["prim"][i]))
            nrm = np.cross(v1 - v0, v2 - v0)
            nrm /= np.linalg.norm(nrm)
            dd = inv[:3, :3] @ rays["direction"][i].astype(np.float64)
            cos = abs(np.dot(nrm, dd)) / max(np.linalg.norm(dd), 1e-30)
            assert err[j] * max(cos, 1e-3) <= tol[j], f"{label}: t err {err[j]} (cos {cos}) > tol {tol[j]} (t={ref['t'][i]})"
        # barycentrics: abs ABS_UV; on small / distant triangles float32 cannot resolve them that finely, so for the few
        # hits over the bound the two barycentric pairs must name the same POINT within 256 ulp of the coordinate scale
        duv = np.maximum(np.abs(gpu["u"][idx] - ref["u"][idx]), np.abs(gpu["v"][idx] - ref["v"][idx]))
        for j in np.nonzero(duv > ABS_UV)[0]:
            i = idx[j]
            tris, inv = scene_lookup(int(ref["inst"][i]))
            v0, v1, v2 = _tri_f64(tris, int(ref["prim"][i]))
            pg = (1 - gpu["u"][i] - gpu["v"][i]) * v0 + gpu["u"][i] * v1 + gpu["v"][i] * v2
            pr = (1 - ref["u"][i] - ref["v"][i]) * v0 + ref["u"][i] * v1 + ref["v"][i] * v2
            oo = inv[:3, :3] @ rays["origin"][i].astype(np.float64) + inv[:3, 3]
            scale = max(np.abs(oo).max(), np.abs(pr).max(), 1e-6)
            bound = 256 * 2.0 ** -24 * scale
            if oracle_artefacts:
                # stress tests: a grazing ray's in-plane hit position is ill-conditioned by 1 / cos(incidence) on BOTH sides
                nrm = np.cross(v1 - v0, v2 - v0)
                dd = inv[:3, :3] @ rays["direction"][i].astype(np.float64)
                cos = abs(np.dot(nrm, dd)) / max(np.linalg.norm(nrm) * np.linalg.norm(dd), 1e-30)
                bound /= max(cos, 1e-4)
            assert np.abs(pg - pr).max() <= bound, f"{label}: barycentric err {duv[j]} = point err {np.abs(pg - pr).max()} at scale {scale}"
    miss = same & (ref["inst"] < 0)
    if miss.any():
        assert np.array_equal(gpu["t"][miss], ref["t"][miss]), f"{label}: miss records must carry tmax"
    bad = np.nonzero(~same)[0]
    assert len(bad) <= max(2, int(max_fraction * n)), f"{label}: {len(bad)} ID mismatches of {n}"
    unexplained = []
    for i in bad:
        o = rays["origin"][i].astype(np.float64)
        d = rays["direction"][i].astype(np.float64)
        cand = []
        for rec in (gpu[i], ref[i]):
            if rec["inst"] < 0:
                cand.append(None)
                continue
            tris, inv = scene_lookup(int(rec["inst"]))
            oo = inv[:3, :3] @ o + inv[:3, 3]
            dd = inv[:3, :3] @ d
            cand.append(exact_hit(oo, dd, *_tri_f64(tris, int(rec["prim"]))))
        ok = False
        ts = [c[0] for c in cand if c is not None]
        if len(ts) == 2 and abs(ts[0] - ts[1]) <= REL_T * max(abs(ts[0]), abs(ts[1]), 1e-6):
            ok = True  # near-tie in t
        for c in cand:
            if c is None:
                continue
            tc, u, v = c
            # a hit within tolerance of the ray interval's own bounds: accepted by one formulation, rejected by the other
            for bound in (float(rays["tmin"][i]), float(rays["tmax"][i])):
                if abs(tc - bound) <= REL_T * max(abs(bound), 1e-6):
                    ok = True
            if min(u, v, 1.0 - u - v) <= EDGE_EPS:
                ok = True  # edge graze: MT and the watertight test may disagree about inside/outside
        if not ok and oracle_artefacts and cand[1] is not None and not (float(rays["tmin"][i]) < cand[1][0] < float(rays["tmax"][i])):
            # the oracle's hit does not exist in exact arithmetic; the product must then report a miss or an exact-valid hit
            ok = cand[0] is None or (float(rays["tmin"][i]) < cand[0][0] < float(rays["tmax"][i]) and min(cand[0][1], cand[0][2], 1.0 - cand[0][1] - cand[0][2]) >= -EDGE_EPS)
        if not ok and oracle_artefacts and cand[1] is None and cand[0] is not None:
            # the oracle's float32 Moller-Trumbore missed a (sub-resolution) triangle that exact arithmetic says the ray hits
            ok = float(rays["tmin"][i]) < cand[0][0] < float(rays["tmax"][i]) and min(cand[0][1], cand[0][2], 1.0 - cand[0][1] - cand[0][2]) >= -EDGE_EPS
        if not ok and oracle_artefacts:
            # a triangle seen EDGE-ON: its projection along the ray is a sliver thinner than the float32 rounding of the vertex coordinates relative to the
            # ray origin, so the sign of an edge function (watertight test) or of a barycentric (Moller-Trumbore) is decided by rounding, either way
            for rec in (gpu[i], ref[i]):
                if rec["inst"] < 0:
                    continue
                tris, inv = scene_lookup(int(rec["inst"]))
                if edge_on(inv[:3, :3] @ o + inv[:3, 3], inv[:3, :3] @ d, *_tri_f64(tris, int(rec["prim"]))):
                    ok = True
        if not ok and reference_epsilon > 0.0 and cand[0] is not None:
            tris, inv = scene_lookup(int(gpu[i]["inst"]))
            det = mt_determinant(inv[:3, :3] @ o + inv[:3, 3], inv[:3, :3] @ d, *_tri_f64(tris, int(gpu[i]["prim"])))
            tg, ug, vg = cand[0]
            valid = float(rays["tmin"][i]) < tg < float(rays["tmax"][i]) and min(ug, vg, 1.0 - ug - vg) >= -EDGE_EPS
            closer = cand[1] is None or tg <= cand[1][0] * (1.0 + REL_T)
            ok = valid and closer and abs(det) < reference_epsilon * 1.001
        if not ok:
            unexplained.append((int(i), gpu[i], ref[i], cand))
    assert not unexplained, f"{label}: unexplained mismatches {unexplained[:3]}"
    return len(bad)


def lookup_from_desc(desc):
    """Builds scene_lookup for a scenes.SceneDesc: global instance id -> (tris, inverse matrix)."""
    table = []
    for mid in sorted(desc.instances):
        mats = np.asarray(desc.instances[mid], np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)
        for M in mats:
            if not M.any() or mid not in desc.meshes:
                table.append(None)
            else:
                table.append((desc.meshes[mid], np.linalg.inv(M)))
    return lambda inst: table[inst]
