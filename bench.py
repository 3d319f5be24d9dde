#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ray-tracing backend.

Metric (BASELINE.json): Mrays/s closest-hit, incoherent rays.  Workload at every N: config C2 of BASELINE.json
(`configs[1]`): synthetic 1M-triangle random soup, 2^24 incoherent random rays, closest hit.
A "step" = one closest-hit pass over the 2^24 rays.  N > 1: the scene is replicated, every rank traces its own
2^24 rays (weak scaling, no data-path collective); value = total rays / max-over-ranks device time.

  value      rays already resident in HBM, device time of the persistent traversal kernel (CUDA events on the
             backend's launching stream, taken inside librfwb200)
  e2e        same metric through the C-ABI host entry point rfwb200_trace_closest with pinned HOST buffers:
             H2D of the rays and D2H of the hit records are inside the timed region, every step
  roofline   HBM roofline of the traversal kernel from ALGORITHMIC bytes (52 B/ray: 32 B ray in + 20 B hit out)
  cpu_baseline  the CPU oracle (a port of the reference path; the reference itself is Rust and cannot be built
             here) on the box's host cores over a bounded prefix of the same rays
  path_tracing  second metric block, at every N: BASELINE.json's "path samples/s" on C5 (configs[4]: 3840x2160, 64 spp, depth 5,
             10 M triangles replicated) tile-sharded over the N ranks, STRONG scaling — the frame's one collective (the NCCL
             accumulator gather inside librfwb200, rfwb200_render_gather) is inside the timed region
  extra      any-hit Mrays/s, BVH build ms, traversal statistics, C3 path tracing (10 k instances, 1080p, 16 spp)

`--impl reference` times the CPU port alone on the same config/metric (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TRIS, SOUP_S, N_RAYS = 1_000_000, 0.005, 1 << 24
BYTES_PER_RAY_CLOSEST = 52
CPU_SAMPLE_RAYS = 1 << 20


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")  # refreshed with every ncu --set full capture of the kernel
    try:
        return float(json.load(open(p))["traffic_bytes_per_launch"]) / 1e9
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region of the default
    run is ~50 ms of kernels, shorter than one `nvidia-smi -lms` period, so the samples come from NVML directly (a thread
    polling every ~2 ms; the traced calls release the GIL); nvidia-smi is the fallback when pynvml cannot be loaded."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0):
        self.index = index
        self.rows = []       # (sm MHz, max MHz, reasons bitmask)
        self.proc = None
        self.nvml = None
        self.handle = None
        self.stop_flag = False
        self.thread = None
        self.source = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._sample_nvml()
            self.rows.clear()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((sm, self.max_mhz, reasons))

    def _poll_nvml(self):
        while not self.stop_flag:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mask = 0
                for (bit, _), v in zip(self.REASONS, f[3:7]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.rows.append((float(f[0]), float(f[1]), mask))
            except ValueError:
                continue

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (pynvml and nvidia-smi unavailable)"], "samples": 0}
        if self.source == "nvml":
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = list(self.rows)
        sm = [r[0] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = sorted(name for bit, name in self.REASONS if mask & bit)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in rows) if rows else None, "reasons": reasons, "samples": len(sm),
                "source": self.source}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def native_oracle_build():
    """The CPU arm's binary: oracle/oracle.cpp compiled HERE, on the box whose host cores are timed, with -march=native (the
    portable liboracle.so that travels with the repository is -march=x86-64-v3).  Same source, same -ffp-contract=off.  Returns a
    description of what will be loaded; falls back to the portable build when there is no compiler."""
    out = os.path.join(ROOT, "oracle", "_native")
    src = os.path.join(ROOT, "oracle", "oracle.cpp")
    try:
        import hashlib

        # keyed by this host's CPU (model + ISA flags): a library built on another machine must never be loaded here
        cpu = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
        lib = os.path.join(out, "liboracle_native_" + hashlib.sha1("".join(cpu).encode()).hexdigest()[:12] + ".so")
        os.makedirs(out, exist_ok=True)
        if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            subprocess.check_call(["/usr/bin/g++", "-O3", "-march=native", "-fopenmp", "-ffp-contract=off", "-fPIC", "-std=c++17", "-shared", "-o", lib, src],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        os.environ["RFWB200_ORACLE_LIB"] = lib
        return "g++ -O3 -march=native -fopenmp -ffp-contract=off, compiled on this host"
    except Exception:
        return "portable build: g++ -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off (no compiler on this host for -march=native)"


def cpu_port_rate(desc, rays, threads=0):
    """The CPU oracle (BVH2 traversal: the faster of its two modes) on a bounded sample; returns Mrays/s."""
    from oracle import oracle as orc

    if threads <= 0:
        threads = host_threads()
    o = orc.OracleBackend(det_eps=0.0, threads=threads)
    desc.apply(o)
    best = None
    for _ in range(2):
        o.trace_closest(rays, mode=orc.MODE_BVH2)
        best = o.trace_seconds if best is None else min(best, o.trace_seconds)
    return len(rays) / best / 1e6, o.max_threads() if threads <= 0 else threads, o.build_seconds


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rfw_rs_b200 import scenes

    desc = scenes.soup_scene(N_TRIS, SOUP_S)
    rays = scenes.random_rays(CPU_SAMPLE_RAYS)
    cpu_build = native_oracle_build()   # (before the oracle module is imported: it picks the library up from the environment)
    from oracle import oracle as orc

    o = orc.OracleBackend(det_eps=0.0, threads=host_threads())
    desc.apply(o)
    for _ in range(args.warmup):
        o.trace_closest(rays[: 1 << 16], mode=orc.MODE_BVH2)
    t = 0.0
    for _ in range(args.steps):
        o.trace_closest(rays, mode=orc.MODE_BVH2)
        t += o.trace_seconds
    value = args.steps * len(rays) / t / 1e6
    sample = f"first {CPU_SAMPLE_RAYS} of the {N_RAYS} rays per step (BVH2 binned-SAH + Moller-Trumbore port, OpenMP)"
    line = {
        "impl": "reference", "metric": "Mrays/s closest-hit (incoherent)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "C2: 1M-triangle random soup, 2^24 incoherent rays, closest hit", "triangles": N_TRIS, "rays_per_step": len(rays)},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": host_threads(), "kind": "port", "sample": sample, "bvh_build_s": o.build_seconds,
                         "per_core": value / max(1, host_threads()), "build": cpu_build,
                         "note": "scalar C++ port of the reference's traversal (binned-SAH BVH2 + Moller-Trumbore, one ray per thread, OpenMP over rays); rtbvh itself has 4-wide packets — a reported baseline, not the optimisation target"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def _render_frames(be, view, spp, depth, frames, torch, dist):
    """`frames` frames of rfwb200_render_gather (render_spp on the owned tiles + the NCCL gather to rank 0), each bracketed by a
    barrier; returns (per-frame ms = max over ranks of the host wall time of the call, stats of the last frame)."""
    times = []
    rs = None
    for _ in range(frames):
        be.reset_accumulator()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        be.render_gather(view, spp, depth, root=0)   # synchronous: returns when this rank's part of the frame (incl. the gather) is done
        rs = be.render_stats()
        t = torch.tensor([rs["frame_ms"]], device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(t.item())
    return times, rs


def _sum_over_ranks(vals, torch, dist):
    t = torch.tensor([float(v) for v in vals], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t)
    return [x.item() for x in t]


def _max_over_ranks(vals, torch, dist):
    t = torch.tensor([float(v) for v in vals], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [x.item() for x in t]


def _ranks_ok(err, torch, dist):
    ok = torch.tensor([0 if err else 1], device="cuda", dtype=torch.int32)
    if dist is not None:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return ok.item() == 1


def path_tracing_block(backend_mod, scenes, sharding, torch, rank, world, dist, frames, n_tris, spp):
    """C5 (BASELINE.json configs[4]): tile-sharded path tracing, strong scaling; the gather is inside the timed region."""
    w, h, depth, tile = 3840, 2160, 5, 64
    err, be = None, None
    uid = sharding.broadcast_unique_id(dist, torch, rank) if dist is not None else None   # (an NCCL id serves ONE communicator)
    try:
        desc = scenes.c5_scene(n_tris)
        be = backend_mod.B200Backend(w, h, device=torch.cuda.current_device(), tile_size=tile, rank=rank, world=world, sky=(0.3, 0.35, 0.5))
        if world > 1:
            be.comm_init(uid, rank, world)
        t0 = time.time()
        desc.apply(be)
        sync_wall_ms = (time.time() - t0) * 1e3
        view = scenes.c5_view(w, h)
    except Exception as ex:
        err = repr(ex)
    if not _ranks_ok(err, torch, dist):
        return {"error": err or "another rank failed"}
    try:
        _render_frames(be, view, spp, depth, 1, torch, dist)        # warm-up frame of the SAME shape: the wave queues are sized by the samples per wave (a 4 spp warm-up left their allocation inside the first timed frame), NCCL buffers, L2
        times, rs = _render_frames(be, view, spp, depth, frames, torch, dist)
    except Exception as ex:
        err = repr(ex)
    if not _ranks_ok(err, torch, dist):
        return {"error": err or "another rank failed"}
    samples, ext, shd = _sum_over_ranks([rs["samples"], rs["extension_rays"], rs["shadow_rays"]], torch, dist)
    render_ms, gather_ms = _max_over_ranks([rs["render_ms"], rs["gather_ms"]], torch, dist)
    bs = be.build_stats()
    cs = _max_over_ranks([bs["checksum"] & 0xFFFFFFFFFFFF], torch, dist)[0], -_max_over_ranks([-(bs["checksum"] & 0xFFFFFFFFFFFF)], torch, dist)[0]
    t_s = (sum(times) / len(times)) / 1e3
    return {
        "metric": "path samples/s (tile-sharded path tracing, accumulator gather over NCCL inside the timed region)", "value": samples / t_s, "unit": "samples/s",
        "Msamples_per_s": samples / t_s / 1e6, "scaling": "strong", "n_gpus": world, "frames_timed": frames, "ms_per_frame": t_s * 1e3, "ms_per_frame_each": times,
        "higher_is_better": True,
        "config": {"workload": f"C5: {n_tris}-triangle soup + ground + 64 area lights (replicated per GPU), 3840x2160, {spp} spp, depth 5, hash RNG, 64x64 tiles in Morton order, tile k -> rank k mod N",
                   "timing": "per frame: barrier, then rfwb200_render_gather (render_spp of the owned tiles + export -> ncclSend/ncclRecv gather on rank 0 -> de-tile + sqrt(acc/spp)); host wall time of the synchronous call, max over ranks, mean over the frames"},
        "render_ms_max_rank": render_ms, "gather_ms_max_rank": gather_ms, "gather_bytes": w * h * 16,
        "extension_rays": ext, "shadow_rays": shd, "Mrays_per_s_all_kinds": (ext + shd) / t_s / 1e6, "mean_segments_per_sample": ext / max(1.0, samples),
        "hbm_roofline_frac_algorithmic": (ext * 320.0 / t_s) / 1e9 / load_peaks()[0] / max(1, world),
        "scene_checksums_equal_across_ranks": bool(cs[0] == cs[1]), "bvh_bytes": int(bs["bvh_bytes"]), "blas_build_ms": bs["blas_build_ms"], "tlas_build_ms": bs["tlas_build_ms"],
        "synchronize_wall_ms_incl_upload": sync_wall_ms, "nccl_version": backend_mod.load_library().rfwb200_nccl_version() if world > 1 else None,
    }


def path_tracing_extra(backend_mod, scenes, sharding, torch, rank, world, spp, dist):
    """C3: 10k-instance scene, 1920x1080, `spp` spp, depth 5 — tile-sharded over the ranks, gather inside the frame time."""
    w, h, depth, tile = 1920, 1080, 5, 64
    err, be = None, None
    uid = sharding.broadcast_unique_id(dist, torch, rank) if dist is not None else None
    try:  # the rank-local part first; a failure here must not leave the other ranks waiting in a collective
        desc = scenes.instanced_scene(grid=100, subdiv=3, n_lights=16)
        be = backend_mod.B200Backend(w, h, device=torch.cuda.current_device(), tile_size=tile, rank=rank, world=world, sky=(0.3, 0.35, 0.5))
        if world > 1:
            be.comm_init(uid, rank, world)
        desc.apply(be)
        view = scenes.camera_view((0.0, 14.0, -62.0), (0.0, -0.25, 1.0), w, h)
    except Exception as ex:
        err = repr(ex)
    if not _ranks_ok(err, torch, dist):
        return {"error": err or "another rank failed"}
    _render_frames(be, view, spp, depth, 1, torch, dist)   # warm-up of the same shape
    times, rs = _render_frames(be, view, spp, depth, 3, torch, dist)
    samples, ext, shd = _sum_over_ranks([rs["samples"], rs["extension_rays"], rs["shadow_rays"]], torch, dist)
    render_ms, gather_ms = _max_over_ranks([rs["render_ms"], rs["gather_ms"]], torch, dist)
    # where the frame goes: one more (untimed) render with events between the stage launches (serialises the stages)
    stage_ms = None
    try:
        be.set_option("stage_timing", 1)
        be.reset_accumulator()
        be.render_spp(view, spp, depth)
        stage_ms = dict(zip(("generate", "extend", "shade", "connect", "reduce_and_bookkeeping"), [round(float(x), 3) for x in be.render_stats()["stage_ms"]]))
        be.set_option("stage_timing", 0)
    except Exception:
        pass
    bs = be.build_stats()
    warm = {"blas_build_ms": None, "tlas_build_ms": None}
    try:  # the numbers above are the first build of this backend (module loading, allocator growth); a rebuild of everything, warm:
        wb, wt = [], []
        for _ in range(3):
            be.set_option("sah_treelet", 8)   # marks every mesh dirty
            be.synchronize()
            b2 = be.build_stats()
            wb.append(b2["blas_build_ms"]); wt.append(b2["tlas_build_ms"])
        warm = {"blas_build_ms": min(wb), "tlas_build_ms": min(wt)}
    except Exception:
        pass
    t_s = min(times) / 1e3
    return {
        "workload": f"C3: 10k icosphere instances (12.8M instanced triangles) + ground + 16 area lights, 1920x1080, {spp} spp, depth 5, tile-sharded; frame = render + NCCL gather (best of 3)",
        "samples_per_s": samples / t_s, "Msamples_per_s": samples / t_s / 1e6, "ms_per_frame": t_s * 1e3, "ms_per_frame_each": times, "render_ms_max_rank": render_ms,
        "gather_ms_max_rank": gather_ms, "extension_rays": ext, "shadow_rays": shd,
        "Mrays_per_s_all_kinds": (ext + shd) / t_s / 1e6, "mean_segments_per_sample": ext / max(1.0, samples),
        "hbm_roofline_frac_algorithmic": (ext * 320.0 / t_s) / 1e9 / load_peaks()[0] / max(1, world),
        "stage_ms_rank0_serialised": stage_ms, "tlas_build_ms_first": bs["tlas_build_ms"], "blas_build_ms_first": bs["blas_build_ms"], "tlas_build_ms": warm["tlas_build_ms"], "blas_build_ms": warm["blas_build_ms"],
        "instances": bs["num_instances"],
    }


def bind_to_gpu_numa_node(index):
    """Host side of the e2e path: run this rank (and first-touch its pinned buffers) on the CPUs NVML reports as local to
    its GPU, so 8 ranks do not push their 832 MB per step across the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        pass  # a hint only


_REAL_STDOUT = None


def _quiet_stdout():
    """The driver reads ONE JSON line from stdout; native libraries (NCCL prints its version banner there) must not
    interleave with it: fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def dynamic_build_extra(backend):
    """Per-frame build costs of dynamic scenes (the fused small / medium builds of builder.cu::k_build_small): warm rebuild of the 170 per-mesh
    BLASes of the reference's pica asset, and N animated CesiumMan instances (4 672 triangles each) re-skinned and rebuilt with a new pose per frame
    (set_skins + synchronize, wall clock).  Fixtures: tests/golden/{pica,cesium_man}.npz (made from /root/reference/assets, committed)."""
    from rfw_rs_b200 import gltf
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    be = backend.B200Backend()
    gltf.per_mesh(gltf.load_npz(os.path.join(gold, "pica.npz"))).apply(be)
    wall = []
    for _ in range(5):
        l0 = be.launch_count()
        be.set_option("sah_treelet", 8)   # marks every mesh dirty
        t0 = time.perf_counter(); be.synchronize(); wall.append((time.perf_counter() - t0) * 1e3)
        launches = be.launch_count() - l0
    # (build_stats once, after the loop: it evaluates the lazy BVH checksum — two launches per mesh)
    out["pica_170_blas_rebuild"] = {"blas_build_ms": be.build_stats()["blas_build_ms"], "synchronize_wall_ms": min(wall[1:]), "kernel_launches": int(launches), "meshes": 170, "triangles": 76274}
    del be
    man = gltf.load_npz(os.path.join(gold, "cesium_man.npz"))
    for copies in (17, 65):
        be = backend.B200Backend()
        gltf.skinned(man, copies=copies).apply(be)
        wall = []
        for f in range(8):
            pose = gltf.pose_joints(man.skins[0], angle=0.1 + 0.03 * f)
            t0 = time.perf_counter(); be.set_skins([pose]); be.synchronize(); wall.append((time.perf_counter() - t0) * 1e3)
        out[f"animated_characters_{copies - 1}"] = {"frame_ms_set_skins_plus_synchronize": min(wall[2:]), "triangles_each": 4672}
        del be
    return out


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-dynamic", action="store_true", help="skip extra.dynamic_scene_builds (pica rebuild, animated characters)")
    ap.add_argument("--pt-spp", type=int, default=16, help="C3 extra: samples per pixel")
    ap.add_argument("--no-path-tracing", action="store_true", help="skip the C5 path-tracing block")
    ap.add_argument("--c5-tris", type=int, default=10_000_000)
    ap.add_argument("--c5-spp", type=int, default=64)
    ap.add_argument("--c5-frames", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from rfw_rs_b200 import backend, scenes, sharding, wire

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        import datetime

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=600))
    # torch.distributed is the plumbing (barriers, max-over-ranks of the timings, distributing the 128-byte NCCL id); the
    # data-path collective — the accumulator gather — is the library's own NCCL communicator (rfwb200_comm_init)

    # ---- scene (replicated) and this rank's rays -------------------------------------------------------
    desc = scenes.soup_scene(N_TRIS, SOUP_S)
    be = backend.B200Backend(device=local_rank)
    for kv in os.environ.get("RFWB200_BENCH_OPTS", "").split(","):  # profiling runs only (e.g. l2_persist=1); the driver's runs set nothing
        if "=" in kv:
            be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    t0 = time.time()
    desc.apply(be)
    sync_wall_ms = (time.time() - t0) * 1e3
    bs = be.build_stats()
    # warm rebuild of the same BLAS (the first build of a process also pays module loading and allocator growth)
    warm_build_ms = []
    for _ in range(3):
        be.set_option("sah_treelet", 8)   # marks every mesh dirty
        be.synchronize()
        warm_build_ms.append(be.build_stats()["blas_build_ms"])
    assert be.build_stats()["checksum"] == bs["checksum"]  # same tree every time
    rays = scenes.random_rays(N_RAYS, start=rank * N_RAYS)
    pin_rays = backend.PinnedArray(N_RAYS, wire.RAY)
    pin_hits = backend.PinnedArray(N_RAYS, wire.HIT)
    pin_rays.array[:] = rays
    d_rays = torch.empty(N_RAYS * 32, dtype=torch.uint8, device="cuda")
    d_hits = torch.empty(N_RAYS * 20, dtype=torch.uint8, device="cuda")
    d_occ = torch.empty(N_RAYS, dtype=torch.int32, device="cuda")
    d_rays.copy_(torch.from_numpy(pin_rays.array.view(np.uint8).reshape(-1)))
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: rays resident in HBM ---------------------------------------------------------------------
    for _ in range(args.warmup):
        be.trace_closest_device(d_rays.data_ptr(), N_RAYS, d_hits.data_ptr())
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = be.launch_count()
    kernel_ms = 0.0
    w0 = time.time()
    for _ in range(args.steps):
        be.trace_closest_device(d_rays.data_ptr(), N_RAYS, d_hits.data_ptr())  # synchronous; CUDA-event time inside
        kernel_ms += be.trace_stats()["kernel_ms"]
    barrier()
    wall_ms = (time.time() - w0) * 1e3
    launches = be.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([kernel_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kernel_ms_max = t.item()
    value = world * args.steps * N_RAYS / (kernel_ms_max / 1e3) / 1e6

    # ---- e2e: host buffers through the C-ABI entry point ----------------------------------------------------
    if os.environ.get("RFWB200_BENCH_STREAMED", "1") == "0":
        # profiler runs only: ncu serialises kernels and copies, so the single persistent launch that consumes rays WHILE
        # they are uploaded can never see its watermark advance; fall back to the chunked copy/launch pipeline there
        be.set_option("streamed", 0)
    for _ in range(max(1, args.warmup - 1)):
        be.trace_closest(pin_rays.array, out=pin_hits.array)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        be.trace_closest(pin_rays.array, out=pin_hits.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * N_RAYS / t.item() / 1e6
    # the same leg with the reference's own 16-byte hit record (rfwb200_trace_closest_packed): 48 instead of 52 bytes per ray on the host link
    pin_packed = backend.PinnedArray(N_RAYS, wire.HIT_PACKED)
    be.trace_closest_packed(pin_rays.array, out=pin_packed.array)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        be.trace_closest_packed(pin_rays.array, out=pin_packed.array)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - e0], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_packed_value = world * args.steps * N_RAYS / t.item() / 1e6
    packed_same = bool(np.array_equal(pin_packed.array["prim"], pin_hits.array["prim"]) and np.array_equal(pin_packed.array["t"], pin_hits.array["t"]))
    pin_packed.free()
    launches += 0  # e2e launches are counted separately below
    hit_rate = float((pin_hits.array["inst"] >= 0).mean())

    # ---- extras -----------------------------------------------------------------------------------------------
    extra = {}
    if rank == 0 or dist is not None:
        be.trace_any_device(d_rays.data_ptr(), N_RAYS, d_occ.data_ptr())
        any_ms = 0.0
        for _ in range(3):
            be.trace_any_device(d_rays.data_ptr(), N_RAYS, d_occ.data_ptr())
            any_ms += be.trace_stats()["kernel_ms"]
        extra["any_hit_Mrays_per_s_per_gpu"] = 3 * N_RAYS / (any_ms / 1e3) / 1e6
    if rank == 0:
        st = be.trace_closest_counted(d_rays.data_ptr(), 1 << 22, d_hits.data_ptr())
        extra["nodes_per_ray"] = st["nodes_visited"] / st["rays"]
        extra["tris_per_ray"] = st["tris_tested"] / st["rays"]
        extra["traversal_bytes_per_ray"] = extra["nodes_per_ray"] * 96 + extra["tris_per_ray"] * 48  # a node visit fetches its whole 96-byte stride (3 sectors)
        # L2-side roofline: measured traversal bytes against a self-measured L2-resident read bandwidth
        l2_peak = be.measure_l2_read_gbs(32 << 20, 50)
        trav_gbs = extra["traversal_bytes_per_ray"] * N_RAYS / (kernel_ms / max(1, args.steps) / 1e3) / 1e9
        extra["l2_roofline"] = {"peak_GBps": l2_peak, "peak_source": "k_l2_read microbenchmark: 32 MiB buffer, L1-bypassing loads, same run",
                                "achieved_GBps": trav_gbs, "frac": trav_gbs / l2_peak if l2_peak else None,
                                "note": "traversal bytes = nodes/ray x 96 B (the stride a visit fetches) + tris/ray x 48 B requested by the SMs, counted by the instrumented kernel in this run; part of them hit in L1 (ncu: profiles/), the rest go to L2"}
        # issue-side roofline: the kernel is bound by warp-instruction issue, not by bytes (ncu: 70 % of the issue slots busy)
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            ipr = prof["warp_instructions_per_launch"] / prof["rays_per_launch"]
            sms = torch.cuda.get_device_properties(0).multi_processor_count
            peak_issue = sms * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6   # 4 schedulers per SM, 1 warp instruction per clock each
            extra["issue_roofline"] = {"warp_instructions_per_ray": ipr, "source": "STATIC, not measured in this run: ncu smsp__inst_executed.sum / rays of the committed capture named in profiles/r2_traffic.json (refreshed with every capture of the kernel); only the rate it is multiplied with is live",
                                       "achieved_Ginst_per_s": ipr * value / max(1, world) * 1e6 / 1e9, "peak_Ginst_per_s": peak_issue / 1e9,
                                       "frac": ipr * value / max(1, world) * 1e6 / peak_issue, "avg_active_threads_of_32": prof["avg_active_threads_per_warp_instruction"]}
        except Exception:
            pass
        extra["hit_rate"] = hit_rate
        extra["bvh_build"] = {"blas_build_ms": min(warm_build_ms), "blas_build_ms_first_in_process": bs["blas_build_ms"], "synchronize_wall_ms_incl_upload": sync_wall_ms, "wide_nodes": bs["blas_nodes"],
                              "bvh_bytes": bs["bvh_bytes"], "sah_cost": bs["sah_cost"]}
    if not args.no_extras and not args.no_dynamic and rank == 0:
        try:
            extra["dynamic_scene_builds"] = dynamic_build_extra(backend)
        except Exception as ex:
            extra["dynamic_scene_builds"] = {"error": repr(ex)}
    if not args.no_extras:
        try:
            pt = path_tracing_extra(backend, scenes, sharding, torch, rank, world, args.pt_spp, dist)
            if rank == 0:
                extra["path_tracing_c3"] = pt
        except Exception as ex:  # the headline metric must still be reported
            if rank == 0:
                extra["path_tracing_c3"] = {"error": repr(ex)}
    # ---- second metric block: C5 tile-sharded path tracing, strong scaling, gather inside the timed region -------------------
    pt_block = None
    if not args.no_path_tracing:
        del be, d_rays, d_hits, d_occ   # the C2 scene and its 1.3 GB of ray / hit buffers are done
        pin_rays.free(); pin_hits.free()
        torch.cuda.empty_cache()
        try:
            pt_block = path_tracing_block(backend, scenes, sharding, torch, rank, world, dist, args.c5_frames, args.c5_tris, args.c5_spp)
        except Exception as ex:
            pt_block = {"error": repr(ex)}

    if rank == 0:
        peak, peak_src = load_peaks()
        per_launch_ms = kernel_ms / max(1, args.steps)
        achieved = BYTES_PER_RAY_CLOSEST * N_RAYS / (per_launch_ms / 1e3) / 1e9
        cpu_build = native_oracle_build() if world == 1 else None
        cpu_rate, cores, cpu_build_s = cpu_port_rate(desc, rays[:CPU_SAMPLE_RAYS]) if world == 1 else (None, None, None)  # the CPU arm is timed at N = 1 only
        line = {
            "metric": "Mrays/s closest-hit (incoherent)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 1M-triangle random soup (s=0.005), 2^24 incoherent rays per GPU, closest hit", "triangles": N_TRIS, "rays_per_step_per_gpu": N_RAYS,
                       "l2_policy": "inputs larger than L2: 512 MiB rays + 320 MiB hits streamed per step (evict-first); the 64 MB BVH is the step's reused working set",
                       "scene": "replicated per GPU", "timing": "CUDA events on the backend stream inside librfwb200, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": N_RAYS * 32, "d2h_bytes_per_step": N_RAYS * 20,
                    "note": "rfwb200_trace_closest with pinned host buffers: ONE persistent launch consumes rays as the upload lands them (device watermark) while completed 2^18-ray granules are downloaded (per-warp progress slots mirrored to the host); wall clock, max over ranks"},
            "e2e_packed_hits": {"value": e2e_packed_value, "unit": "Mrays/s", "h2d_bytes_per_step": N_RAYS * 32, "d2h_bytes_per_step": N_RAYS * 16, "same_hits_as_e2e": packed_same,
                                "note": "rfwb200_trace_closest_packed: the reference's own 16-byte hit record (inst, prim, t, 2 x 16-bit barycentrics; ray_extend.comp:267) instead of the 20-byte RfwHit"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(), "traffic_unit": "GB per launch (ncu dram__bytes_read+write, profiles/r2_traffic.json)",
                         "algorithmic_GB_per_launch": BYTES_PER_RAY_CLOSEST * N_RAYS / 1e9,
                         "kernel": "k_trace_persistent<RayBufferIO, closest, single-level>", "algorithmic_bytes_per_ray": BYTES_PER_RAY_CLOSEST, "peak_source": peak_src,
                         "note": "pointer-chasing traversal over an L2-resident BVH: the HBM fraction is small by construction (SURVEY 8d); see extra.traversal_bytes_per_ray for the L2-side traffic"},
            "cpu_baseline": None if cpu_rate is None else {"value": cpu_rate, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"first {CPU_SAMPLE_RAYS} rays of rank 0's step (best of 2); oracle BVH2 binned-SAH + Moller-Trumbore, OpenMP", "bvh_build_s": cpu_build_s,
                             "per_core": cpu_rate / max(1, cores),
                             "build": cpu_build,
                             "note": "scalar C++ port (rtbvh itself has 4-wide packets); "
                                     "pinned bit for bit against the reference's own shaders compiled for the host (oracle/_ref, tests/test_ref_glsl.py), whose 1e-4 determinant epsilon "
                                     "would reject nearly every triangle of this soup — the timed arm uses epsilon 0 and returns the GPU's hits. A reported baseline, not the optimisation target."},
            "wall_ms_timed_region": wall_ms,
            "path_tracing": pt_block,
            "extra": extra,
        }
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
