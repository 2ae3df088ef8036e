"""The N > 1 path on hardware: a 2-rank torchrun launch over NCCL (skipped on a one-GPU box) of tests/nccl_worker.py --
point split, bucket-range split, error propagation through the exchange, one Groth16 proof over both GPUs -- every
result compared with the oracle on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29713", os.path.join(ROOT, "tests", "nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"NCCL_WORKER_OK world={world}" in out.stdout
