"""Launches tests/dist_check.py on the visible GPUs (>= 2) with torchrun: the sharded fwd+bwd must
equal the single-graph fp64 oracle, both with the NVLS multicast exchange fused into the producer
kernels and with plain NCCL all-gathers.  Skipped on a single-GPU box; `gpurun --gpus 2` and the
driver's multi-GPU runs exercise it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("comm", ["pull", "multicast", "nccl"])
def test_sharded_step_matches_oracle(comm):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ, HAN_DIST_COMM=comm)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + ["pull", "multicast", "nccl"].index(comm)),
           os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("comm", ["pull", "nccl"])
def test_tile_sharded_step_with_other_exchanges(comm):
    """The same tile-sharded check with the Z re-sharding as an NCCL all-to-all (HAN_DIST_COMM=nccl) and with
    copy-engine pulls inside a split meta-path's ranks (HAN_TILE_COMM=pull)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ, HAN_DIST_PARTITION="tile")
    env.update({"HAN_DIST_COMM": "nccl"} if comm == "nccl" else {"HAN_TILE_COMM": "pull"})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29543 + (comm == "nccl")), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_tile_sharded_step_matches_oracle():
    """(meta-path x row-block) tile sharding (han_b200/tiles.py): 2 ranks own two / one whole meta-paths each;
    4 ranks additionally split a meta-path into two row blocks (the T / R exchange runs inside that pair)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    # HAN_PUSH_MIN_ROWS=32: the chunked "push" exchange of a split meta-path runs in several chunks even on these
    # few-hundred-row graphs
    env = dict(os.environ, HAN_DIST_PARTITION="tile", HAN_PUSH_MIN_ROWS="32")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29539", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
