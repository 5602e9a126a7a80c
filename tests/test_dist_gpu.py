"""On-device data-parallel parity: after one train step from identical weights on different
shards, every rank holds identical parameters, equal to a 1-GPU step on the concatenated batch
(scripts/dp_check.py under torchrun, NCCL).  Needs two visible GPUs; skips itself on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_step_equals_single_gpu_step_on_the_concatenated_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 visible GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
