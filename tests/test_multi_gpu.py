"""GPU tier, needs >= 2 GPUs: the NCCL paths (structure sharding + slab halo merge) under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_torchrun_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(here, "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
