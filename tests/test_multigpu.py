"""NCCL spectral sharding on >= 2 GPUs (skipped on a single-GPU box): every rank runs
its shard with full-band control flow, one all-gather closes the run, and the result is
bit-identical to the single-GPU run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nspec", [100, 33])
def test_sharded_equals_single_gpu(nspec):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multigpu_worker.py"),
           str(nspec)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTIGPU_OK" in r.stdout


def test_raman_owner_shards_equal_single_gpu():
    """rt_run_rrs_sharded (owner ranges + locally recomputed elastic halo + one all-gather) == single-GPU rt_run_rrs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "multigpu_worker.py"), "300", "rrs"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTIGPU_OK" in r.stdout
