"""Tiled schedule over NCCL on two real GPUs (skipped on a single-GPU box): tools/tiled_run.py under
torchrun, phase 1 per rank, all-gather exchange, joined rounds, checked against the tiled CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_gpu_tiled_run_matches_tiled_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "tiled_run.py"), "1536", "1024", "8", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "partition identical: True" in r.stdout, r.stdout[-2000:]
