"""GPU, >= 2 devices: sharded ensembles / large-N force / large-N ham_soft equal their single-GPU results
(tools/check_mgpu.py under torchrun over NCCL).  Skipped on a one-GPU box; the N > 1 host logic is also covered
on CPU by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_two_gpus_match_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "check_mgpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=850)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
