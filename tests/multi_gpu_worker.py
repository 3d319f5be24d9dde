"""One rank of tests/test_multi_gpu.py: a process per GPU, no torch.distributed — the ranks exchange the 128-byte NCCL unique
id through a file, everything else goes through the C ABI (rfwb200_comm_init / rfwb200_render_gather / rfwb200_gather_image)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rfw_rs_b200 import backend, scenes  # noqa: E402


def main():
    rank, world, workdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    w, h, spp, depth = 320, 192, 4, 4   # 5 x 3 tiles of 64: ranks own different tile counts (padding path)
    id_path = os.path.join(workdir, "nccl_id.bin")
    if rank == 0:
        uid = backend.B200Backend.comm_unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    desc = scenes.instanced_scene(grid=8, subdiv=2, n_lights=4)
    view = scenes.camera_view((0, 3.5, -9.0), (0, -0.35, 1.0), w, h)
    be = backend.B200Backend(w, h, device=rank, sky=(0.3, 0.35, 0.5))   # created as a single rank ...
    be.comm_init(uid, rank, world)                                      # ... the communicator sets the tile sharding
    desc.apply(be)
    be.render_gather(view, spp, depth, root=0)
    rs = be.render_stats()
    if rank == 0:
        np.save(os.path.join(workdir, "gathered_root.npy"), be.read_output())
    be.reset_accumulator()
    be.render_spp(view, spp, depth)
    be.gather_image(root=0xFFFFFFFF)                                    # all-gather: every rank holds the frame
    np.save(os.path.join(workdir, f"gathered_all_{rank}.npy"), be.read_output())
    np.save(os.path.join(workdir, f"stats_{rank}.npy"), np.array([rs["samples"], rs["gather_ms"], rs["frame_ms"], rs["render_ms"]], np.float64))
    if rank == 0:  # the same frame on one GPU
        one = backend.B200Backend(w, h, device=0, sky=(0.3, 0.35, 0.5))
        desc.apply(one)
        one.render_spp(view, spp, depth)
        np.save(os.path.join(workdir, "single.npy"), one.read_output())
    be.comm_destroy()


if __name__ == "__main__":
    main()
