"""The CUDA path of the edge-cut partitioned processor on 2 GPUs of one box (``-m gpu``; skipped with fewer than 2 visible GPUs):
tests/partition_gpu_worker.py under torch.distributed.run with NCCL -- the NVLink peer-memory halo against the NCCL halo and against
the single-GPU run of the same step.  (The gloo tests in tests/test_partition_cpu.py cover the bookkeeping without a GPU.)"""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_processor_peer_halo_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "partition_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and run.stdout.count("PARTITION-GPU-OK") == world, run.stdout[-3000:] + run.stderr[-5000:]
    print("\n" + "\n".join(ln for ln in run.stdout.splitlines() if "PARTITION-GPU-OK" in ln))
