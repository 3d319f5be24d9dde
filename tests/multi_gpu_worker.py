"""One rank of tests/test_multi_gpu.py: a process per GPU, no torch.distributed — the ranks exchange the 128-byte NCCL unique
id through a file, everything else goes through the C ABI (rfwb200_comm_init / rfwb200_render_gather / rfwb200_gather_image)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rfw_rs_b200 import backend, scenes  # noqa: E402


def deadline_case(rank, world, workdir, uid):
    """Rank 1 never joins the gather: rank 0 must come back with an error after its deadline instead of hanging (ncclCommAbort)."""
    w, h = 128, 64
    desc = scenes.instanced_scene(grid=4, subdiv=1, n_lights=2)
    view = scenes.camera_view((0, 3.0, -7.0), (0, -0.4, 1.0), w, h)
    be = backend.B200Backend(w, h, device=rank)
    be.comm_init(uid, rank, world)
    desc.apply(be)
    be.render_spp(view, 1, 2)
    if rank == 0:
        be.set_option("gather_timeout_s", 3)
        t0 = time.time()
        try:
            be.gather_image(root=0)
            msg = "gather returned"
        except backend.RfwError as e:
            msg = str(e)
        with open(os.path.join(workdir, "deadline.txt"), "w") as f:
            f.write(f"{time.time() - t0:.2f}\n{msg}\n")
        be.render_spp(view, 1, 2)          # the backend is still usable on its own
        with open(os.path.join(workdir, "deadline_after.txt"), "w") as f:
            f.write(str(be.sample_count))
    else:
        time.sleep(8.0)
    sys.stdout.flush()
    os._exit(0)                            # (no orderly NCCL teardown with an aborted peer)


def main():
    rank, world, workdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    mode = sys.argv[4] if len(sys.argv) > 4 else "gather"
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
    if mode == "deadline":
        return deadline_case(rank, world, workdir, uid)
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
