"""Host-side plumbing of the multi-GPU path: one process per GPU, scene replicated, work sharded, ONE collective.

Path tracing shards by image tile (64x64, Morton order, tile k -> rank k mod n); ray batches shard
by contiguous range.  The only collective is the final accumulator gather: NCCL inside librfwb200 (`rfwb200_comm_init` /
`rfwb200_gather_image`); the CPU tests exercise the same tile layout with gloo and the numpy mirror `assemble_host`.  The tile layout itself comes from the library
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


def broadcast_unique_id(dist, torch, rank):
    """Distributes rank 0's ncclGetUniqueId bytes over an existing torch.distributed group (any backend): the bootstrap of the
    library's own communicator.  A host without torch distributes the 128 bytes by any other means (tests/multi_gpu_worker.py
    uses a file)."""
    uid = backend.B200Backend.comm_unique_id() if rank == 0 else bytes(128)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def gather_image(be, root=0, d_image_ptr=None):
    """The frame's one collective, inside the library: rfwb200_gather_image (export tiles -> NCCL -> de-tile + sqrt(acc / spp)),
    all on the backend's own stream.  The receiver's image lands in its output buffer (be.read_output()) or in d_image_ptr."""
    be.gather_image(root, d_image_ptr)
