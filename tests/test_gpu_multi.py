"""Multi-GPU (>= 2 visible devices): sharded rollouts under torchrun/NCCL.
Skipped on a single-GPU box; the host-side logic is covered on CPU by
tests/test_distributed_cpu.py (gloo, world size 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_rollouts_two_gpus():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("MULTI_GPU_RESULT ")][-1]
    out = json.loads(line[len("MULTI_GPU_RESULT "):])
    assert out["private_traces_equal_oracle"]
    assert out["episodes"][0] == out["episodes"][1]
    assert out["sum_return_close"] and out["max_return_equal"]
    assert out["shared_replicas_identical"] and out["shared_learned"]
