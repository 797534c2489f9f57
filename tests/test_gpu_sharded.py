"""GPU test of the sharded path on ONE device: two ranks (two processes, both on cuda:0) each own a
shard of the actors; the rank-sum exchange goes over gloo.  Checks the CUDA path + ngu_b200.sharding
against the unsharded oracle.  (With >= 2 GPUs the same script runs over NCCL: tools/multi_gpu_check.py.)"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_two_shards_one_gpu_gloo_exchange(cuda_device):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--backend", "gloo", "--same-gpu"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("sharded parity ok") == 2
