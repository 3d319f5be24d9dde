#!/usr/bin/env python
"""Runs the BASELINE.json configs that bench.py does not headline (C1, C4, C5) and prints one JSON line per config.

  C1  primary-ray closest hit of the glTF assets (fixtures in tests/golden) at 1280x720, variants C1a / C1b
  C4  5M-triangle soup + 256 emissive triangles: closest hits of 2^24 rays, then one NEE any-hit ray per hit
  C5  3840x2160, 64 spp, depth 5, 10M-triangle soup + ground + 64 area lights, tile-sharded over the ranks
      (launch with torchrun for N > 1; the final gather is rfwb200_gather_image: NCCL inside the library)

usage: python scripts/run_configs.py --configs c1,c4,c5 [--c5-spp 64] [--c5-tris 10000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rfw_rs_b200 import backend, gltf, scenes, sharding, wire  # noqa: E402


def dev(arr):
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).cuda()


def run_c1(out):
    gold = os.path.join(ROOT, "tests", "golden")
    for name in ("cesium_man", "pica"):
        asset = gltf.load_npz(os.path.join(gold, name + ".npz"))
        for label, desc in (("C1a flattened", gltf.flatten(asset)), ("C1b per-mesh BLAS + TLAS", gltf.per_mesh(asset))):
            w, h = 1280, 720
            be = backend.B200Backend(w, h)
            t0 = time.time(); desc.apply(be); sync_ms = (time.time() - t0) * 1e3
            view = gltf.c1_camera(gltf.flatten(asset), w, h)
            best = 1e9
            for _ in range(5):
                hits = be.cast_primary(view)
                best = min(best, be.trace_stats()["kernel_ms"])
            bs = be.build_stats()
            out({"config": "C1", "asset": name, "variant": label, "triangles": int(bs["num_triangles"]), "instances": int(bs["num_instances"]), "rays": w * h,
                 "generate+trace_ms": best, "Mrays_per_s": w * h / best / 1e3, "hit_fraction": float((hits["inst"] >= 0).mean()),
                 "blas_build_ms": bs["blas_build_ms"], "tlas_build_ms": bs["tlas_build_ms"], "synchronize_wall_ms": sync_ms, "sah_cost": bs["sah_cost"]})


def run_c4(out, n_tris, n_rays):
    desc = scenes.c4_scene(n_tris)
    be = backend.B200Backend()
    t0 = time.time(); desc.apply(be); sync_ms = (time.time() - t0) * 1e3
    bs = be.build_stats()
    rays = scenes.random_rays(n_rays)
    d_rays = dev(rays)
    d_hits = torch.empty(n_rays * 20, dtype=torch.uint8, device="cuda")
    best_c = 1e9
    for _ in range(3):
        be.trace_closest_device(d_rays.data_ptr(), n_rays, d_hits.data_ptr())
        best_c = min(best_c, be.trace_stats()["kernel_ms"])
    hits = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=wire.HIT)
    sh, _ = scenes.c4_shadow_rays(desc, rays, hits)
    L = desc.area_lights
    d_sh = dev(sh)
    d_occ = torch.empty(len(sh), dtype=torch.int32, device="cuda")
    best_a = 1e9
    for _ in range(3):
        be.trace_any_device(d_sh.data_ptr(), len(sh), d_occ.data_ptr())
        best_a = min(best_a, be.trace_stats()["kernel_ms"])
    occ = d_occ.cpu().numpy()
    out({"config": "C4", "triangles": int(bs["num_triangles"]), "area_lights": len(L), "closest_rays": n_rays, "closest_Mrays_per_s": n_rays / best_c / 1e3,
         "shadow_rays": len(sh), "any_hit_Mrays_per_s": len(sh) / best_a / 1e3, "any_hit_ms": best_a, "unoccluded_fraction": float((occ == 0).mean()),
         "hbm_roofline_frac_any_hit": (len(sh) * 36 / (best_a / 1e3)) / 1e9 / 6453.1, "blas_build_ms": bs["blas_build_ms"], "synchronize_wall_ms": sync_ms,
         "bvh_bytes": int(bs["bvh_bytes"])})


def run_c5(out, n_tris, spp, rank, world, dist):
    w, h, depth, tile = 3840, 2160, 5, 64
    desc = scenes.c5_scene(n_tris)
    be = backend.B200Backend(w, h, device=torch.cuda.current_device(), tile_size=tile, rank=rank, world=world, sky=(0.3, 0.35, 0.5))
    t0 = time.time(); desc.apply(be); sync_ms = (time.time() - t0) * 1e3
    bs = be.build_stats()
    view = scenes.c5_view(w, h)
    if dist is not None:
        be.comm_init(sharding.broadcast_unique_id(dist, torch, rank), rank, world)
    be.render_spp(view, 1, depth)
    be.reset_accumulator()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    be.render_spp(view, spp, depth)
    rs = be.render_stats()
    ms = torch.tensor([rs["render_ms"]], device="cuda")
    tot = torch.tensor([float(rs["samples"]), float(rs["extension_rays"]), float(rs["shadow_rays"])], device="cuda", dtype=torch.float64)
    cs = torch.tensor([bs["checksum"] & 0x7FFFFFFFFFFFFFFF], device="cuda", dtype=torch.int64)
    cs_min, cs_max = cs.clone(), cs.clone()
    gather_ms = 0.0
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
        dist.all_reduce(cs_min, op=dist.ReduceOp.MIN); dist.all_reduce(cs_max, op=dist.ReduceOp.MAX)
        be.gather_image(0)  # warm-up of the communicator and the buffers
        dist.barrier()
        be.gather_image(0)  # export tiles -> NCCL gather to rank 0 -> assemble + sqrt(acc/spp), inside librfwb200
        gather_ms = be.render_stats()["gather_ms"]
    if rank == 0:
        t_s = ms.item() / 1e3
        out({"config": "C5", "n_gpus": world, "triangles": int(bs["num_triangles"]), "resolution": [w, h], "spp": spp, "depth": depth,
             "samples": tot[0].item(), "render_ms": ms.item(), "Msamples_per_s": tot[0].item() / t_s / 1e6, "extension_rays": tot[1].item(), "shadow_rays": tot[2].item(),
             "Mrays_per_s_all_kinds": (tot[1].item() + tot[2].item()) / t_s / 1e6, "mean_segments_per_sample": tot[1].item() / tot[0].item(),
             "gather_ms": gather_ms, "gather_bytes": w * h * 16, "scene_checksums_equal_across_ranks": bool(cs_min.item() == cs_max.item()),
             "blas_build_ms": bs["blas_build_ms"], "synchronize_wall_ms": sync_ms, "bvh_bytes": int(bs["bvh_bytes"])})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c4,c5")
    ap.add_argument("--c4-tris", type=int, default=5_000_000)
    ap.add_argument("--c4-rays", type=int, default=1 << 24)
    ap.add_argument("--c5-tris", type=int, default=10_000_000)
    ap.add_argument("--c5-spp", type=int, default=64)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import datetime

        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=300))

    def out(d):
        print(json.dumps(d), flush=True)

    cfgs = args.configs.split(",")
    if "c1" in cfgs and rank == 0:
        run_c1(out)
    if "c4" in cfgs and rank == 0:
        run_c4(out, args.c4_tris, args.c4_rays)
    if "c5" in cfgs:
        run_c5(out, args.c5_tris, args.c5_spp, rank, world, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
