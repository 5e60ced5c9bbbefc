"""Data-parallel step on real GPUs (NCCL over NVLink): gradients of the sharded step equal the single-process big-batch
gradients and are identical on every rank, eager and in CUDA-graph mode (tools/dp_check.py under torchrun).
Skipped on boxes with fewer than two GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_two_rank_nccl_gradients_match_big_batch():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical across ranks = True" in r.stdout
