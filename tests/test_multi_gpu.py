"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a 1-GPU box): single-image query sharding
(dagl_ce_forward_rows_f32 + one NCCL all-gather + dagl_ce_fold_rows_f32) reproduces the 1-GPU forward."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_query_sharding_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SHARDED_CHECK OK" in res.stdout
