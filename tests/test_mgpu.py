"""Multi-GPU parity under pytest: launches tests/mgpu_check.py (z-slab results against the whole-grid CPU oracle) on two ranks.
Skips cleanly on a box with fewer than two GPUs."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_two_slab_parity_against_the_whole_grid_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-6000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "ALL PASS" in p.stdout and "FAIL " not in p.stdout
