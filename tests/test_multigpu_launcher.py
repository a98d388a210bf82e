"""Runs tests/run_multigpu_parity.py (one rank per GPU, NCCL inside the C ABI) under `pytest -m gpu` whenever the box
has at least two GPUs, so that multi-GPU parity is part of the recorded GPU suite and not a side script."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multigpu_parity_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (this box has {n}); run tests/run_multigpu_parity.py under torchrun")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_multigpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert f"multi-GPU parity OK: world={world} profile=compact" in r.stdout
