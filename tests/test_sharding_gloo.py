"""CPU tier, world_size 2 and 3 over gloo: the host logic of the N > 1 path — tile ownership (Morton order, tile k ->
rank k mod n), the tile-major gather layout and its reassembly, and the contiguous ray-range split.  The GPU side of the
same layout (export / assemble kernels) is checked by tests/test_gpu_parity.py::test_tile_sharding_is_invariant."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rfw_rs_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        order, (tx, ty) = sharding.tile_layout(w, h, tile)
        mine = sharding.owned_tiles(w, h, tile, rank, world)
        tpr = sharding.tiles_per_rank(w, h, tile, world)
        # every rank "renders" its tiles: pixel value = global pixel id (+ rank in channel 1), tile-major like export_tiles
        send = np.zeros((tpr, tile, tile, 4), np.float32)
        for k, t in enumerate(mine):
            x0, y0 = (int(t) % tx) * tile, (int(t) // tx) * tile
            ys, xs = np.mgrid[y0:y0 + tile, x0:x0 + tile]
            inside = (xs < w) & (ys < h)
            send[k, ..., 0] = np.where(inside, xs + ys * w, 0)
            send[k, ..., 1] = np.where(inside, rank + 1, 0)
        recv = torch.empty(world * send.size, dtype=torch.float32)
        dist.all_gather_into_tensor(recv, torch.from_numpy(send.reshape(-1)))
        counts = torch.tensor([len(mine)], dtype=torch.int64)
        dist.all_reduce(counts)
        b, e = sharding.ray_range(1000003, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (b, e))
        if rank == 0:
            img = sharding.assemble_host(recv.numpy().reshape(-1, 4), w, h, tile, world, tpr)
            expect = (np.arange(w * h, dtype=np.float32)).reshape(h, w)
            ok_pixels = bool(np.array_equal(img[..., 0], expect))
            owners = img[..., 1]
            ok_owner = bool(owners.min() >= 1 and owners.max() <= world)
            ok_count = int(counts.item()) == len(order) == tx * ty
            spans.sort()
            ok_spans = spans[0][0] == 0 and spans[-1][1] == 1000003 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            balance = max(e - b for b, e in spans) - min(e - b for b, e in spans)
            q.put((ok_pixels, ok_owner, ok_count, ok_spans, balance, len(set(order.tolist())) == len(order)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h,tile", [(2, 200, 120, 32), (3, 130, 70, 16), (2, 64, 64, 64)])
def test_tile_gather_layout_over_gloo(world, w, h, tile):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, tile, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ok_pixels, ok_owner, ok_count, ok_spans, balance, unique = q.get(timeout=10)
    assert ok_pixels and ok_owner and ok_count and ok_spans and unique
    assert balance <= 1


def test_morton_order_is_a_z_curve():
    order, (tx, ty) = sharding.tile_layout(256, 256, 64)
    assert (tx, ty) == (4, 4)
    # Z-order of a 4x4 grid (x fastest inside each 2x2 block)
    assert order.tolist() == [0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15]
    # tile k -> rank k mod n spreads neighbouring tiles over the ranks
    assert sharding.owned_tiles(256, 256, 64, 1, 4).tolist() == [1, 3, 9, 11]
