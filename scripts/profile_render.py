"""C3 wavefront render for profiling: env SPP (default 4), DEPTH (5), W/H (1920x1080), GRID (100)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfw_rs_b200 import backend, scenes
w, h = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
spp, depth, grid = int(os.environ.get("SPP", 4)), int(os.environ.get("DEPTH", 5)), int(os.environ.get("GRID", 100))
c5 = os.environ.get("SCENE", "").startswith("c5:")   # SCENE=c5:<triangles>: the C5 soup from the C5 camera (use W=3840 H=2160)
desc = scenes.c5_scene(int(os.environ["SCENE"][3:])) if c5 else scenes.instanced_scene(grid=grid, subdiv=3, n_lights=16)
be = backend.B200Backend(w, h, sky=(0.3, 0.35, 0.5), tile_size=64, rank=int(os.environ.get("RANK_", 0)), world=int(os.environ.get("WORLD_", 1))); desc.apply(be)
for kv in os.environ.get("OPTS", "").split(","):
    if "=" in kv: be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
view = scenes.c5_view(w, h) if c5 else scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, h)
if "WAVE_PATHS" in os.environ: be.set_option("wave_paths", int(os.environ["WAVE_PATHS"]))
be.render_spp(view, 1, depth); be.reset_accumulator()
for _ in range(int(os.environ.get("REPS", 2))):
    be.reset_accumulator(); be.render_spp(view, spp, depth)
    rs = be.render_stats(); print(rs, "Msamples/s", rs["samples"] / rs["render_ms"] / 1e3, "Mrays/s", (rs["extension_rays"] + rs["shadow_rays"]) / rs["render_ms"] / 1e3)
if os.environ.get("STAGES", "1") == "1":
    be.set_option("stage_timing", 1); be.reset_accumulator(); be.render_spp(view, spp, depth); rs = be.render_stats()
    print("stage ms [generate, extend, shade, connect, reduce+bookkeeping]:", [round(x, 3) for x in rs["stage_ms"]], "sum", round(sum(rs["stage_ms"]), 3), "render_ms", round(rs["render_ms"], 3))
    be.set_option("stage_timing", 0)
if os.environ.get("L2", "0") == "1": print("l2 read GB/s", be.measure_l2_read_gbs(32 << 20, 50), "64MB:", be.measure_l2_read_gbs(64 << 20, 30), "512MB (HBM):", be.measure_l2_read_gbs(512 << 20, 5))
