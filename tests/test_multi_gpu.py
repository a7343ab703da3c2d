"""Multi-GPU scheduler on real GPUs (needs >= 2 devices; skipped otherwise): the NVLink
peer-memory halo push against the NCCL exchange and against a single-GPU evaluation."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("cfg,n_per", [("C2", 64000), ("C4", 96000), ("C5", 128000)])
def test_peer_push_equals_nccl_and_single_gpu(cfg, n_per):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "mg", "peer_vs_nccl.py"), cfg, str(n_per)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "PEER_VS_NCCL_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
