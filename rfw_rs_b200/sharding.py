"""Host-side plumbing of the multi-GPU path: one process per GPU, scene replicated, work sharded, ONE collective.

Path tracing shards by image tile (64x64, Morton order, tile k -> rank k mod n); ray batches shard
by contiguous range.  The only collective is the final accumulator gather (`torch.distributed.all_gather_into_tensor`
over NCCL on the GPU box, gloo in the CPU tests).  The tile layout itself comes from the library
(`rfwb200_tile_layout`, host-only) so this file, the export kernel and the assemble kernel cannot disagree.
"""
import numpy as np

from . import backend


def tile_layout(width, height, tile=64):
    """All tile ids in Morton order (numpy uint32) and the tile grid size."""
    L = backend.load_library()
    n = L.rfwb200_tile_layout(width, height, tile, None, 0)
    out = np.zeros(n, dtype=np.uint32)
    L.rfwb200_tile_layout(width, height, tile, out.ctypes.data, n)
    return out, ((width + tile - 1) // tile, (height + tile - 1) // tile)


def owned_tiles(width, height, tile, rank, world):
    order, _ = tile_layout(width, height, tile)
    return order[rank::world]


def tiles_per_rank(width, height, tile, world):
    order, _ = tile_layout(width, height, tile)
    return (len(order) + world - 1) // world


def ray_range(n_rays, rank, world):
    """Contiguous [begin, end) share of a ray batch (C2/C4 style sharding)."""
    base, rem = divmod(n_rays, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def assemble_host(gathered, width, height, tile, world, tiles_per_rank_):
    """numpy mirror of k_wf_assemble without the sqrt: gathered[(rank, local tile, y, x, c)] -> image[h, w, c]."""
    order, (tx, _) = tile_layout(width, height, tile)
    c = gathered.shape[-1]
    g = gathered.reshape(world, tiles_per_rank_, tile, tile, c)
    img = np.zeros((height, width, c), dtype=gathered.dtype)
    for mr, t in enumerate(order):
        r, tl = mr % world, mr // world
        x0, y0 = (int(t) % tx) * tile, (int(t) // tx) * tile
        x1, y1 = min(width, x0 + tile), min(height, y0 + tile)
        img[y0:y1, x0:x1] = g[r, tl, : y1 - y0, : x1 - x0]
    return img


def gather_image(be, dist, torch, width, height, tile, world):
    """Device-side gather used by bench.py: export this rank's tiles, all_gather, assemble + sqrt(acc/spp) on the device."""
    tpr = be.tiles_per_rank
    send = torch.zeros(tpr * tile * tile * 4, dtype=torch.float32, device="cuda")
    recv = torch.empty(world * tpr * tile * tile * 4, dtype=torch.float32, device="cuda")
    image = torch.empty(height * width * 4, dtype=torch.float32, device="cuda")
    be.export_tiles_device(send.data_ptr(), tpr)
    dist.all_gather_into_tensor(recv, send)
    be.assemble_tiles_device(recv.data_ptr(), tpr, world, image.data_ptr())
    return image.view(height, width, 4)
