"""GPU, >= 2 devices: the multi-GPU paths over NCCL (lineage-sharded binning + all-reduce, chain sharding, tempered
ladders across ranks) against the single-device path.  Skipped on a one-GPU box; tests/test_distributed_cpu.py covers the
same host logic over gloo."""
import os
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def test_multi_gpu_paths_match_single_device():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(REPO, "tests", "_multigpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTIGPU OK world=%d" % world in p.stdout
