"""GPU tests that need >= 2 devices (skipped on a one-GPU box): the NVLink p2p exchange that bench.py times at
N > 1, checked against the unsharded oracle, and the learner gradient all-reduce hook behind the reference's own
R2D2 network.  On one GPU the same data plane is covered by tests/test_gpu_sharded.py (gloo exchange) and by the
in-run `parity` object of every multi-GPU bench.py line."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, timeout=600):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.timeout(600)
def test_p2p_exchange_vs_unsharded_oracle(cuda_device):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    res = _torchrun(n, "tools/multi_gpu_check.py", "--exchange", "p2p", "--actors", "37", "--capacity", "3000", "--steps", "10")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("sharded parity ok (p2p)") == n


@pytest.mark.timeout(900)
def test_learner_allreduce_matches_full_batch_on_nccl(cuda_device):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref (the staged reference network) is not present")
    res = _torchrun(min(n, 8), "tools/learner_dp_check.py", timeout=880)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('{"learner_dp_check"')][-1]
    out = json.loads(line)["learner_dp_check"]
    for key in ("blocking", "overlap", "p2p_blocking", "p2p_overlap"):     # NCCL and the library's own NVLink kernel
        assert out[key]["worst_rel_vs_full_batch"] < 2e-4 and out[key]["identical_across_ranks"]
        assert out[key]["copies_in_steady_state"] == 0


def test_learner_hook_single_gpu_is_a_noop_and_keeps_views(cuda_device):
    import torch
    import torch.nn as nn
    from ngu_b200.learner_allreduce import GradAllReducer
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 8)).cuda()
    ps = list(m.parameters())
    red = GradAllReducer(ps, overlap=True)           # overlap requested: silently off for one process
    x = torch.randn(16, 64, device="cuda")
    red.zero_grad()
    m(x).pow(2).mean().backward()
    want = [p.grad.clone() for p in ps]
    red(ps)
    for p, w in zip(ps, want):
        assert torch.equal(p.grad, w)
    assert all(p.grad.data_ptr() == v.data_ptr() for _, bps, vs in red.buckets for p, v in zip(bps, vs))
    assert red.copies_in == 0
