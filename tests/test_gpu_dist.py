"""Launches tests/dist_check.py on every visible GPU (>= 2) with torchrun: sharded fwd+bwd over NCCL
must equal the single-graph fp64 oracle.  Skipped on a single-GPU box; the driver's multi-GPU runs
and `gpurun --gpus 2` exercise it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_step_matches_oracle():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
