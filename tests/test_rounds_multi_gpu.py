"""Sharded rounds on real GPUs (NCCL) against the unsharded run: identical accepted records and cell
tables for every GPU count (tier 4).  Needs >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_rounds_equal_single_gpu(built_lib, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "rounds_multi_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ROUNDS_MULTI OK" in res.stdout
