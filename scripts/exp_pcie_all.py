"""Aggregate host<->device copy bandwidth with every rank copying at once (torchrun): the ceiling of bench.py's e2e at N GPUs.
Each rank: 512 MiB pinned -> device and 320 MiB device -> pinned on two streams, the sizes of one e2e step."""
import os, time
import torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hr = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); hh = torch.empty(320 << 20, dtype=torch.uint8).pin_memory()
dr = torch.empty(512 << 20, dtype=torch.uint8, device="cuda"); dh = torch.empty(320 << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def step(h2d, d2h):
    if h2d:
        with torch.cuda.stream(s1): dr.copy_(hr, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2): hh.copy_(dh, non_blocking=True)
def run(h2d, d2h, reps=5):
    step(h2d, d2h); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): step(h2d, d2h)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / reps
for name, a, b in (("H2D 512 MiB", 1, 0), ("D2H 320 MiB", 0, 1), ("both", 1, 1)):
    dt = run(a, b)
    gb = ((512 << 20) * a + (320 << 20) * b) * world / 1e9
    if rank == 0: print(f"{world} ranks, {name} per rank at once: {dt * 1e3:.2f} ms per step (max over ranks) = {gb / dt:.1f} GB/s aggregate; e2e ceiling {world * (1 << 24) / dt / 1e6:.0f} Mrays/s", flush=True)
if world > 1: dist.destroy_process_group()
