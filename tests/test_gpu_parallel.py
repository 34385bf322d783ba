"""Data-parallel step on real GPUs (needs >= 2): tests/dp_worker.py under torchrun -- shards + global
denominators + the library's symmetric-memory all-reduce kernel reproduce the single-GPU gradients,
eagerly and from a replayed CUDA graph; tools/ar_check.py checks the kernel alone against NCCL."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _torchrun(script, n, timeout):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_step_matches_single_gpu():
    assert "dp_worker OK" in _torchrun("tests/dp_worker.py", 2, 600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allreduce_kernel_matches_nccl():
    assert "ar_check OK" in _torchrun("tools/ar_check.py", 2, 600)
