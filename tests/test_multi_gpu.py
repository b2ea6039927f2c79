"""The in-kernel dataset gather on two GPUs of one box (skipped on single-GPU boxes): every rank must end
up with exactly the dataset the NCCL all_gather produces (tests/multi_gpu_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_in_kernel_gather_equals_nccl_gather():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("multi_gpu_check ok") == 8, r.stdout[-2000:]      # six gather cases (both forms) + two host-level checks
