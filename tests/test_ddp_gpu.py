"""Launches tests/ddp_parity_2gpu.py under torchrun when two CUDA devices are visible (gpurun --gpus 2)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_gpu_gradient_exchange_equals_dense_allreduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run through gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(ROOT / "tests" / "ddp_parity_2gpu.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DDP PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
