"""material textures + equirect skybox.  CPU tier: known answers for the oracle's samplers
(shade.comp:268-281, :90-96; sampler state backends/gpu-rt/src/lib.rs:1026-1034, :471-480).  GPU tier: the CUDA
shade stage against the oracle on a textured scene."""
import numpy as np
import pytest

from rfw_rs_b200 import scenes


def _rgb(tex, level, x, y):
    """texel (x, y) of a mip level as float RGB, honouring the BGRA/RGBA byte order"""
    px = tex.levels[level][y, x].astype(np.float32) / 255.0
    return px[[2, 1, 0]] if tex.format == 0 else px[:3]


def test_oracle_sampler_known_answers(oracle_mod):
    cpu = oracle_mod.OracleBackend()
    bgra = scenes.pattern_texture(16, 1, "checker", fmt=0)
    rgba = scenes.pattern_texture(16, 1, "checker", fmt=1)
    cpu.set_textures([bgra, rgba])
    assert bgra.mip_levels == 5 and [l.shape[0] for l in bgra.levels] == [16, 8, 4, 2, 1]
    w = 16
    for (x, y) in [(0, 0), (5, 9), (15, 15)]:
        u, v = (x + 0.5) / w, (y + 0.5) / w
        # level 0 is bilinear (mag filter Linear): exact texel value at a texel centre
        np.testing.assert_allclose(cpu.sample_texture(0, 0, u, v, 0.0)[:3], _rgb(bgra, 0, x, y), atol=1e-6)
        # same picture through the other byte order
        np.testing.assert_allclose(cpu.sample_texture(1, 0, u, v, 0.0)[:3], cpu.sample_texture(0, 0, u, v, 0.0)[:3], atol=1e-6)
        # Repeat addressing
        np.testing.assert_allclose(cpu.sample_texture(0, 0, u + 3.0, v - 2.0, 0.0), cpu.sample_texture(0, 0, u, v, 0.0), atol=1e-5)
        # levels >= 1 are nearest (min filter Nearest)
        np.testing.assert_allclose(cpu.sample_texture(0, 0, u, v, 1.0)[:3], _rgb(bgra, 1, x // 2, y // 2), atol=1e-6)
        np.testing.assert_allclose(cpu.sample_texture(0, 0, u, v, 2.0)[:3], _rgb(bgra, 2, x // 4, y // 4), atol=1e-6)
        # LOD beyond the last level clamps
        np.testing.assert_allclose(cpu.sample_texture(0, 0, u, v, 9.0)[:3], _rgb(bgra, 4, 0, 0), atol=1e-6)
    # bilinear halfway between two texel centres = their mean
    a, b = _rgb(bgra, 0, 3, 4), _rgb(bgra, 0, 4, 4)
    np.testing.assert_allclose(cpu.sample_texture(0, 0, 4.0 / w, 4.5 / w, 0.0)[:3], 0.5 * (a + b), atol=1e-6)
    # fetchTexelTrilinear: (1-f) * level0 + f * level1 with level0 = int(lambda), f = fract(lambda)
    u, v = 5.5 / w, 9.5 / w
    l1, l2 = cpu.sample_texture(0, 0, u, v, 1.0), cpu.sample_texture(0, 0, u, v, 2.0)
    np.testing.assert_allclose(cpu.sample_texture(0, 1, u, v, 1.25), 0.75 * l1 + 0.25 * l2, atol=1e-6)
    # the reference's quirk for negative lambda: int() truncates to level 0 but fract = lambda - floor(lambda)
    l0 = cpu.sample_texture(0, 0, u, v, 0.0)
    np.testing.assert_allclose(cpu.sample_texture(0, 1, u, v, -0.25), 0.25 * l0 + 0.75 * l1, atol=1e-6)


def test_oracle_skybox_lookup(oracle_mod):
    cpu = oracle_mod.OracleBackend()
    sky = scenes.pattern_texture(32, 2, "sky", fmt=0)
    cpu.set_skybox(sky)
    # ClampToEdge + bilinear on every level: beyond the border the edge texel repeats
    np.testing.assert_allclose(cpu.sample_texture(-1, 2, -0.3, 0.5, 0.0), cpu.sample_texture(-1, 2, 0.0, 0.5, 0.0), atol=1e-6)
    np.testing.assert_allclose(cpu.sample_texture(-1, 2, 0.5 / 64, 0.5 / 32, 0.0)[:3], _rgb(sky, 0, 0, 0), atol=1e-6)
    np.testing.assert_allclose(cpu.sample_texture(-1, 2, 1.5 / 32, 1.5 / 16, 1.0)[:3], _rgb(sky, 1, 1, 1), atol=1e-6)


def test_oracle_constant_skybox_equals_constant_sky(oracle_mod):
    """A one-colour skybox must give the image the constant-sky path gives (same paths, same RNG)."""
    desc = scenes.instanced_scene(grid=3, subdiv=1, n_lights=2)
    w, h = 48, 32
    view = scenes.camera_view((0, 2.0, -5.0), (0, -0.3, 1.0), w, h)
    a = oracle_mod.OracleBackend(); desc.apply(a)
    ref, _ = a.render(view, w, h, 2, 3, sky=(51 / 255.0, 102 / 255.0, 153 / 255.0))
    img = np.zeros((8, 16, 4), np.uint8); img[...] = (153, 102, 51, 255)  # BGRA
    desc.skybox = scenes.Texture(img, 3, 0)
    b = oracle_mod.OracleBackend(); desc.apply(b)
    got, _ = b.render(view, w, h, 2, 3, sky=(9, 9, 9))
    np.testing.assert_allclose(got, ref, atol=2e-6)


def test_oracle_textured_scene_differs_from_untextured(oracle_mod):
    desc = scenes.textured_scene(grid=2, subdiv=1, tex_size=16)
    w, h = 40, 30
    view = scenes.camera_view((0, 2.0, -4.0), (0, -0.35, 1.0), w, h)
    a = oracle_mod.OracleBackend(); desc.apply(a)
    tex, _ = a.render(view, w, h, 2, 3)
    plain = scenes.textured_scene(grid=2, subdiv=1, tex_size=16)
    plain.materials["flags"] = 0
    b = oracle_mod.OracleBackend(); plain.apply(b)
    ref, _ = b.render(view, w, h, 2, 3)
    assert np.isfinite(tex).all() and tex[..., :3].mean() > 0.01
    assert np.abs(tex - ref).mean() > 1e-3  # the maps are actually sampled


@pytest.mark.gpu
def test_gpu_textured_scene_matches_oracle(oracle_mod):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfw_rs_b200 import backend as B
    from tests.test_gpu_parity import check_image

    desc = scenes.textured_scene(grid=4, subdiv=2, tex_size=64)
    w, h, spp, depth = 192, 108, 8, 4
    view = scenes.camera_view((0, 2.5, -6.0), (0, -0.3, 1.0), w, h)
    gpu = B.B200Backend(w, h, sky=(9, 9, 9)); desc.apply(gpu)
    cpu = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(cpu)
    gpu.render_spp(view, spp, depth)
    acc = gpu.read_accumulator()
    ref, _ = cpu.render(view, w, h, spp, depth, clamp=10.0, sky=(9, 9, 9))
    assert ref[..., :3].mean() > 0.05
    full, trimmed = check_image(acc / spp, ref / spp, "textured scene")
    # the textures matter: the same scene with the map flags cleared is a different image
    plain = scenes.textured_scene(grid=4, subdiv=2, tex_size=64); plain.materials["flags"] = 0; plain.skybox = None
    g2 = B.B200Backend(w, h, sky=(0.3, 0.3, 0.3)); plain.apply(g2)
    g2.render_spp(view, spp, depth)
    assert np.abs(g2.read_accumulator() / spp - acc / spp).mean() > 1e-2
    # replacing one texture through `changed` re-uploads only that slot; clearing the skybox falls back to the constant
    desc.textures[2] = scenes.pattern_texture(32, 77, "checker", fmt=1)
    gpu.set_textures(desc.textures, changed=[0, 0, 1]); gpu.set_skybox(None); gpu.synchronize()
    gpu.reset_accumulator(); gpu.render_spp(view, spp, depth)
    cpu.set_textures(desc.textures); cpu.set_skybox(None)
    ref2, _ = cpu.render(view, w, h, spp, depth, clamp=10.0, sky=(9, 9, 9))
    check_image(gpu.read_accumulator() / spp, ref2 / spp, "after texture update")


def test_oracle_debug_views_known_answers(oracle_mod):
    """RenderMode Normal / Albedo / GBuffer at the primary hit: a ground quad facing +y seen from above."""
    sc = scenes.SceneDesc()
    g = scenes.quad((0, 0, 0), (0, 1, 0), 4.0, 4.0, mat_id=0)
    if g["normal"][0, 1] < 0:
        g = scenes.make_triangles(g["vertex0"], g["vertex2"], g["vertex1"], 0)
    sc.meshes[0] = g
    sc.instances[0] = scenes.to_column_major([scenes.trs((0, 1.5, 0))])
    sc.materials = scenes.material(color=(0.2, 0.4, 0.6))
    o = oracle_mod.OracleBackend(); sc.apply(o)
    w, h = 32, 24
    view = scenes.camera_view((0, 5.0, 0.0), (0, -1.0, 1e-3), w, h)
    n = o.debug_view(view, w, h, 1)[h // 2, w // 2]
    a = o.debug_view(view, w, h, 2)[h // 2, w // 2]
    p = o.debug_view(view, w, h, 3)[h // 2, w // 2]
    np.testing.assert_allclose(n[:3], (0, 1, 0), atol=1e-6)
    np.testing.assert_allclose(a, (0.2, 0.4, 0.6, 0.0), atol=1e-6)
    assert abs(p[1] - 1.5) < 1e-4 and abs(p[3] - 3.5) < 1e-2   # hit on the plane y = 1.5, 3.5 below the camera


@pytest.mark.gpu
def test_gpu_debug_views_match_oracle(oracle_mod):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rfw_rs_b200 import backend as B

    desc = scenes.textured_scene(grid=4, subdiv=2, tex_size=64)
    w, h = 160, 90
    view = scenes.camera_view((0, 2.5, -6.0), (0, -0.3, 1.0), w, h)
    gpu = B.B200Backend(w, h); desc.apply(gpu)
    cpu = oracle_mod.OracleBackend(det_eps=0.0); desc.apply(cpu)
    gpu.render_spp(view, 2, 3)
    acc_before = gpu.read_accumulator().copy()
    for mode, tol in ((1, 2e-3), (2, 2e-3), (3, 2e-3)):
        gpu.render(None, view, mode)
        got = gpu.read_output()
        ref = cpu.debug_view(view, w, h, mode)
        # silhouette pixels may resolve to the neighbouring surface (float32 near-ties): compare all but a handful
        bad = (np.abs(got - ref).max(axis=2) > tol * np.maximum(1.0, np.abs(ref).max(axis=2)))
        assert bad.mean() < 5e-3, (mode, bad.mean())
        assert np.abs(ref[..., :3]).sum() > 0
    assert np.array_equal(gpu.read_accumulator(), acc_before)   # debug views leave the accumulation alone
    assert gpu.sample_count == 2
