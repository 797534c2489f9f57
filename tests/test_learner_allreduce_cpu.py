"""world_size-2 gloo test of the learner gradient all-reduce hook."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 8), nn.ReLU(), nn.Linear(8, 4))


def _worker(rank, world, port, out_dir, bucket_bytes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngu_b200.learner_allreduce import GradAllReducer
    model = _model()
    params = list(model.parameters())
    red = GradAllReducer(params, bucket_bytes=bucket_bytes)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    loss = ((model(xs) - ys) ** 2).mean()
    loss.backward()
    red(params)                                   # the seam: between backward and clip
    nn.utils.clip_grad_norm_(params, 40)
    torch.save([p.grad.clone() for p in params], os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("bucket_bytes", [1 << 20, 600])
def test_gradients_are_averaged_over_ranks(tmp_path, bucket_bytes):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), bucket_bytes), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    model = _model()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    ((model(x) - y) ** 2).mean().backward()      # full batch on one process == mean of the two half-batch grads
    for a, b, p in zip(g0, g1, model.parameters()):
        assert torch.equal(a, b)
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)


def test_single_process_is_a_noop():
    from ngu_b200.learner_allreduce import GradAllReducer
    m = _model()
    ps = list(m.parameters())
    red = GradAllReducer(ps)
    m(torch.ones(2, 16)).sum().backward()
    before = [p.grad.clone() for p in ps]
    red(ps)
    assert all(torch.equal(a, p.grad) for a, p in zip(before, ps))
    assert red.nbytes == sum(p.numel() * 4 for p in ps)
