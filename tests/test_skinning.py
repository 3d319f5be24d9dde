"""skinned instances.  SkinnedTriangles3D::apply (crates/rfw-backend/src/structs.rs:820-877) restated in
the oracle and run on the device (k_skin_triangles + a BLAS per skinned instance).  CPU tier: known answers of the
oracle's skinning on the CesiumMan fixture (real JOINTS_0 / WEIGHTS_0 / inverse bind matrices, synthetic pose)."""
import os

import numpy as np
import pytest

from rfw_rs_b200 import gltf, scenes, wire

HERE = os.path.dirname(os.path.abspath(__file__))


def _asset():
    return gltf.load_npz(os.path.join(HERE, "golden", "cesium_man.npz"))


def test_fixture_carries_skin_data():
    a = _asset()
    m = a.meshes[0]
    assert m["joints"].shape == (len(m["positions"]), 4) and m["weights"].shape == m["joints"].shape
    assert np.allclose(m["weights"].sum(axis=1), 1.0, atol=1e-6)
    assert len(a.skins) == 1 and a.skins[0].shape[1:] == (4, 4) and m["joints"].max() < len(a.skins[0])
    jd = gltf.joint_data(m)
    assert jd.dtype.itemsize == 32 and len(jd) == 3 * len(m["indices"])


def test_oracle_skinning_known_answers(oracle_mod):
    a = _asset()
    nj = len(a.skins[0])
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(-1), (nj, 1))
    # identity joint matrices: the skinned copy equals the bind pose up to the rounding of sum(w) * I
    sc = gltf.skinned(a, pose=ident)
    o = oracle_mod.OracleBackend(); sc.apply(o)
    sk = o.skinned_triangles(0, 0, wire.RT_TRIANGLE)
    assert len(sk) == len(sc.meshes[0])
    for f in ("vertex0", "vertex1", "vertex2", "n0"):
        np.testing.assert_allclose(sk[f], sc.meshes[0][f], atol=2e-6)
    np.testing.assert_allclose(sk["normal"], sc.meshes[0]["normal"], atol=1e-4)  # recomputed in float32 from the vertices
    # one rigid transform on every joint moves the whole mesh rigidly: vertices by M, normals by M^-T
    ang = 0.7
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = (0.3, -0.2, 0.5)
    rigid = np.tile(M.T.astype(np.float32).reshape(-1), (nj, 1))
    sc2 = gltf.skinned(a, pose=rigid)
    o2 = oracle_mod.OracleBackend(); sc2.apply(o2)
    sk2 = o2.skinned_triangles(0, 0, wire.RT_TRIANGLE)
    want = sc.meshes[0]["vertex1"].astype(np.float64) @ R.T + M[:3, 3]
    np.testing.assert_allclose(sk2["vertex1"], want, atol=1e-5)
    np.testing.assert_allclose(sk2["n1"], sc.meshes[0]["n1"].astype(np.float64) @ R.T, atol=1e-5)
    np.testing.assert_allclose(sk2["normal"], sc.meshes[0]["normal"].astype(np.float64) @ R.T, atol=1e-4)
    # tracing the rigidly skinned mesh == tracing the bind pose with the inverse-transformed rays
    rays = scenes.random_rays(4000, lo=-1.0, hi=1.0)
    h2 = o2.trace_closest(rays)
    back = rays.copy()
    N = np.asarray(sc.instances[0], np.float64).reshape(-1, 4, 4)[0].T   # the instance (node) matrix: world = N * M_skin * v
    Mi = np.linalg.inv(N @ M @ np.linalg.inv(N))
    back["origin"] = (rays["origin"].astype(np.float64) @ Mi[:3, :3].T + Mi[:3, 3]).astype(np.float32)
    back["direction"] = (rays["direction"].astype(np.float64) @ Mi[:3, :3].T).astype(np.float32)
    h1 = o.trace_closest(back)
    agree = (h1["prim"] == h2["prim"]).mean()
    assert agree > 0.995 and (h2["prim"] >= 0).sum() > 20
    both = (h1["prim"] == h2["prim"]) & (h1["prim"] >= 0)
    np.testing.assert_allclose(h1["t"][both], h2["t"][both], rtol=2e-4)
    # a real pose deforms: different from the bind pose, bounded displacement
    sc3 = gltf.skinned(a)
    o3 = oracle_mod.OracleBackend(); sc3.apply(o3)
    sk3 = o3.skinned_triangles(0, 0, wire.RT_TRIANGLE)
    d = np.linalg.norm(sk3["vertex0"] - sc.meshes[0]["vertex0"], axis=1)
    assert d.max() > 1e-3 and d.max() < 1.0 and np.isfinite(sk3["normal"]).all()
    # an instance with skin id -1 keeps the mesh's own geometry
    sc4 = gltf.skinned(a, copies=3)
    o4 = oracle_mod.OracleBackend(); sc4.apply(o4)
    assert len(o4.skinned_triangles(0, 0, wire.RT_TRIANGLE)) and len(o4.skinned_triangles(0, 1, wire.RT_TRIANGLE))
    assert len(o4.skinned_triangles(0, 2, wire.RT_TRIANGLE)) == 0


@pytest.mark.gpu
def test_gpu_skinned_instances_match_oracle(oracle_mod):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfw_rs_b200 import backend as B
    from tests import parity

    a = _asset()
    sc = gltf.skinned(a, copies=3)   # two skinned copies + one bind-pose instance of the same mesh
    gpu = B.B200Backend(); sc.apply(gpu)
    cpu = oracle_mod.OracleBackend(det_eps=0.0); sc.apply(cpu)
    view = gltf.c1_camera(sc, 320, 180)
    rays = cpu.primary_rays(view, 320, 180)

    def lookup_for(o):
        base = parity.lookup_from_desc(sc)
        def look(inst):
            rec = base(inst)
            if rec is None:
                return None
            sk = o.skinned_triangles(0, inst, wire.RT_TRIANGLE)   # single skinned mesh with id 0: instance index == global id
            return (sk if len(sk) else rec[0], rec[1])
        return look

    ref = cpu.trace_closest(rays)
    assert (ref["inst"] >= 0).sum() > 500 and len(set(ref["inst"][ref["inst"] >= 0])) == 3
    parity.compare_hits(rays, gpu.trace_closest(rays), ref, lookup_for(cpu), "skinned cesium man")
    before = gpu.build_stats()
    # a new pose: only the skinned copies are rebuilt
    sc.skins = [gltf.pose_joints(a.skins[0], angle=0.6, seed=11)]
    gpu.set_skins(sc.skins); gpu.synchronize()
    cpu.set_skins(sc.skins); cpu.synchronize()
    ref2 = cpu.trace_closest(rays)
    assert (ref2["prim"] != ref["prim"]).mean() > 0.005
    parity.compare_hits(rays, gpu.trace_closest(rays), ref2, lookup_for(cpu), "skinned cesium man, second pose")
    assert gpu.build_stats()["blas_nodes"] == before["blas_nodes"]  # the mesh's own BLAS was not rebuilt
    # dropping the skin ids returns every instance to the bind pose
    sc.instance_skins[0][:] = -1
    gpu.set_3d_instances(0, sc.instances[0], skin_ids=sc.instance_skins[0]); gpu.synchronize()
    cpu.set_3d_instances(0, sc.instances[0], skin_ids=sc.instance_skins[0]); cpu.synchronize()
    parity.compare_hits(rays, gpu.trace_closest(rays), cpu.trace_closest(rays), parity.lookup_from_desc(sc), "bind pose again")


@pytest.mark.gpu
def test_animated_crowd_rebuilds_match_the_general_builder(oracle_mod):
    """Per-frame re-skinning of a crowd (builder job scheduler: deformed buffers kept from frame to frame, old BLASes back to the stream-ordered pool,
    the 4 672-triangle characters built by one 512-thread CTA each in the fused launch): over several poses the hits are bit-identical to a backend
    that sends every build through the general builder, and the last frame matches the oracle."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfw_rs_b200 import backend as B
    from tests import parity

    a = _asset()
    sc = gltf.skinned(a, copies=9)   # eight skinned copies + one in the bind pose
    fused = B.B200Backend(); sc.apply(fused)
    general = B.B200Backend(); general.set_option("build_fused", 0); sc.apply(general)
    view = gltf.c1_camera(sc, 640, 240)
    cpu = oracle_mod.OracleBackend(det_eps=0.0); sc.apply(cpu)
    rays = cpu.primary_rays(view, 640, 240)
    last = None
    for frame in range(4):
        pose = [gltf.pose_joints(a.skins[0], angle=0.15 + 0.1 * frame, seed=5 + frame)]
        l0 = fused.launch_count(); fused.set_skins(pose); fused.synchronize(); l_fused = fused.launch_count() - l0
        l0 = general.launch_count(); general.set_skins(pose); general.synchronize(); l_general = general.launch_count() - l0
        hf, hg = fused.trace_closest(rays), general.trace_closest(rays)
        assert hf.tobytes() == hg.tobytes(), frame
        assert (hf["inst"] >= 0).sum() > 200
        if last is not None:
            assert (hf["prim"] != last["prim"]).mean() > 0.002   # the crowd moved
        last = hf
        assert l_fused * 8 < l_general, (l_fused, l_general)   # 8 characters: a handful of launches against ~40 per character
    cpu.set_skins(pose); cpu.synchronize()

    def look(inst):
        rec = parity.lookup_from_desc(sc)(inst)
        if rec is None:
            return None
        sk = cpu.skinned_triangles(0, inst, wire.RT_TRIANGLE)
        return (sk if len(sk) else rec[0], rec[1])

    parity.compare_hits(rays, last, cpu.trace_closest(rays), look, "animated crowd, last frame")
    assert fused.build_stats()["checksum"] == general.build_stats()["checksum"]
