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


def _worker_steps(rank, world, port, out_dir, mode):
    """Three optimiser steps through the seam; `mode` = how the caller zeroes gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngu_b200.learner_allreduce import GradAllReducer
    model = _model()
    params = list(model.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    red = GradAllReducer(params, bucket_bytes=600)
    g = torch.Generator().manual_seed(9)
    for it in range(3):
        x = torch.randn(8, 16, generator=g)
        y = torch.randn(8, 4, generator=g)
        xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
        if mode == "reducer":
            red.zero_grad()
        else:
            opt.zero_grad()                       # set_to_none: the bucket views are dropped and must be re-adopted
        ((model(xs) - ys) ** 2).mean().backward()
        red(params)
        opt.step()
    torch.save({"w": [p.detach().clone() for p in params], "copies_in": red.copies_in,
                "views": all(p.grad.data_ptr() == v.data_ptr() for _, ps, vs in red.buckets for p, v in zip(ps, vs))},
               os.path.join(out_dir, f"s{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("mode", ["reducer", "optimizer"])
def test_training_steps_match_full_batch(tmp_path, mode):
    """Gradients live in the flat bucket (views): three data-parallel SGD steps equal three full-batch steps, whether
    the caller zeroes through the reducer (no copies at all) or through the optimizer (grads re-adopted)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_steps, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    s0, s1 = torch.load(tmp_path / "s0.pt"), torch.load(tmp_path / "s1.pt")
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(9)
    for it in range(3):
        x = torch.randn(8, 16, generator=g)
        y = torch.randn(8, 4, generator=g)
        opt.zero_grad()
        ((model(x) - y) ** 2).mean().backward()
        opt.step()
    for a, b, p in zip(s0["w"], s1["w"], model.parameters()):
        assert torch.equal(a, b)
        assert torch.allclose(a, p.detach(), rtol=1e-5, atol=1e-6)
    assert s0["views"] and s1["views"]
    assert (s0["copies_in"] == 0) == (mode == "reducer")


def test_r2d2_seam_patch_single_process():
    """attach_r2d2_allreduce on an R2D2Learner-shaped object: same update as the reference statements."""
    from ngu_b200.learner_allreduce import attach_r2d2_allreduce
    from ngu_b200.running_mean_std import RunningMeanStd

    class L:
        def __init__(self):
            self.policy = _model()
            self.optimizer = torch.optim.SGD(self.policy.parameters(), lr=0.05)
            self.model_hypr = {"adam_clip_norm": 40}
            self.update_count = 0
            self.r2d2_loss_rms = RunningMeanStd(use_mpi=False)
            self.logged = []
            self.logger = self

        def log_scalar(self, name, v, step):
            self.logged.append((name, float(v), step))

    a, b = L(), L()
    attach_r2d2_allreduce(a, overlap=False)
    x = torch.ones(5, 16)
    for _ in range(2):
        td = a.policy(x)                                   # stands in for the TD errors (trace, batch)
        a.step(td, torch.ones(4))
        td = b.policy(x)
        loss = (td.pow(2).mean(dim=0) * torch.ones(4)).mean()
        b.optimizer.zero_grad(); loss.backward()
        nn.utils.clip_grad_norm_(b.policy.parameters(), 40); b.optimizer.step()
    for p, q in zip(a.policy.parameters(), b.policy.parameters()):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)
    assert a.update_count == 2 and a.logged[-1][0] == "R2D2Loss"


def test_ngu_ar_header_symbols_are_exported():
    """include/ngu_ar.h (the library's own NVLink all-reduce transport): every declared entry point is exported and bound."""
    import os
    import re
    from ngu_b200 import _cabi
    from tests.conftest import ROOT
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "ngu_ar.h")).read()
    declared = set(re.findall(r"\b(ngu_ar_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.AR_SYMBOLS), declared ^ set(_cabi.AR_SYMBOLS)
    for name in declared:
        getattr(lib, name)
    assert int(re.search(r"#define NGU_AR_HANDLE_BYTES (\d+)", header).group(1)) == _cabi.AR_HANDLE_BYTES
