"""-m gpu, needs two GPUs (gpurun --gpus 2): the multi-GPU path behind the C ABI — replicated scene, tile sharding set by
rfwb200_comm_init, ONE collective (the accumulator gather over NCCL inside librfwb200).  The assembled frame must equal the
single-GPU frame bit for bit (RNG streams are keyed by the global pixel id; the gather only moves tiles)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_gather_image_over_nccl_matches_single_gpu(tmp_path):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    world = 2
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_gpu_worker.py"), str(r), str(world), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), outs
    single = np.load(tmp_path / "single.npy")
    root = np.load(tmp_path / "gathered_root.npy")
    assert single[..., :3].mean() > 0.05
    assert np.array_equal(root, single)                       # gather to rank 0 (ncclSend / ncclRecv)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"gathered_all_{r}.npy"), single)   # all-gather: every rank holds the frame
    st = [np.load(tmp_path / f"stats_{r}.npy") for r in range(world)]
    assert sum(s[0] for s in st) == 320 * 192 * 4             # the ranks' samples partition the frame
    # gather_ms measured (the first gather of a communicator includes NCCL's lazy connection set-up: up to seconds); frame_ms covers render + gather
    assert all(0.0 < s[1] < 20000.0 and s[2] >= 0.9 * s[3] for s in st), [list(s) for s in st]


@pytest.mark.gpu
def test_gather_with_a_missing_peer_returns_an_error_not_a_hang(tmp_path):
    """A rank that never joins the frame's collective (crashed, or a host bug) must not hang the others: rfwb200_gather_image waits with
    a deadline (option gather_timeout_s, default 120 s; 3 s here), aborts the communicator (ncclCommAbort) and returns an error; the backend
    stays usable."""
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_gpu_worker.py"), str(r), "2", str(tmp_path), "deadline"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=120)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), outs
    secs, msg = open(tmp_path / "deadline.txt").read().split("\n")[:2]
    assert "no answer from the other ranks" in msg, (msg, outs)
    assert 2.5 <= float(secs) < 20.0
    assert open(tmp_path / "deadline_after.txt").read() == "2"
