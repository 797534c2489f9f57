"""Learner all-reduce hook on real hardware: correctness + overlap behind a real backward.

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/learner_dp_check.py

The network is the REFERENCE's own R2D2 ``DuelingLSTM`` (ngu/models/r2d2/dueling_lstm.py, 8.19 M parameters,
imported from oracle/_ref - test infrastructure) inside the reference's own ``R2D2Learner``, patched at its
seam by ``ngu_b200.learner_allreduce.attach_r2d2_allreduce`` (r2d2_learner.py:53-54).

Checks (every rank):
  1. data-parallel gradient == full-batch gradient: each rank back-propagates its slice of a batch through a
     T-step LSTM unroll; after the reducer the averaged gradient must equal the gradient of the whole batch
     computed on one GPU - for overlap=False and overlap=True (side stream + post-accumulate hooks);
  2. the gradients live in the flat buckets (no copies in the steady state);
  3. timing: backward alone | backward + blocking all-reduce | backward + overlapped all-reduce, and the raw
     bus bandwidth of the 32.75 MB bucket.
Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


class _L:
    def log_scalar(self, *a, **k):
        pass


def build_learner(n_act=18, batch=64):
    from oracle.ref_runner import import_reference
    _, _, model_hypr, ptu = import_reference(use_gpu=True)
    from ngu.models.r2d2.r2d2_learner import R2D2Learner
    torch.manual_seed(1234)                          # identical weights on every rank
    hy = dict(model_hypr, batch_size=batch)
    learner = R2D2Learner(n_act, (1, 84, 84), None, hy, _L())
    learner.to(ptu.device)
    return learner, hy


def unroll_td(policy, obs, act, rew_i, rew_e, beta, T):
    """A T-step unroll of the reference network; returns a (T, B) stand-in for the TD errors."""
    B = obs.shape[1]
    policy.reset_hidden_state(B)
    out = []
    for t in range(T):
        q = policy(obs[t], act[t], rew_i[t], rew_e[t], beta)          # (B, n_act)
        out.append(q.gather(1, act[t].view(B, 1)).squeeze(1) - 0.5)
    return torch.stack(out)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    cuda = torch.cuda.is_available()              # (CPU + gloo: a dry run of the script's logic in the build container)
    if cuda:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    sync = torch.cuda.synchronize if cuda else (lambda: None)
    # the check compares a sliced batch with the full batch to 2e-4: keep the convolutions / GEMMs in full float32
    # (TF32's 10-bit mantissa alone moves a gradient by ~1e-3); the timing section below switches it back on
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ngu_b200.learner_allreduce import GradAllReducer, attach_r2d2_allreduce

    T, Bg = 8, 8 * world                               # global batch, sliced over the ranks
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(T, Bg, 1, 84, 84, generator=g).to(dev)
    # one action per time step for the whole batch: the reference's make_one_hot_batch (pytorch_util.py:52-56) sets the
    # columns of EVERY row from all indices of the batch, so with per-sample actions a slice of the batch would not see
    # the inputs the full batch sees
    act = (torch.arange(T) % 18).view(T, 1).expand(T, Bg).contiguous().to(dev)
    rew_i = torch.randn(T, Bg, 1, generator=g).to(dev)
    rew_e = torch.randn(T, Bg, 1, generator=g).to(dev)
    beta = torch.zeros(Bg, 32, device=dev)
    beta[:, 3] = 1.0
    w = torch.ones(Bg, device=dev)
    lo, hi = rank * Bg // world, (rank + 1) * Bg // world
    sl = lambda x: x[:, lo:hi]

    # full-batch gradient on every rank (same weights, same data): the thing to reproduce
    learner, hy = build_learner()
    params = list(learner.policy.parameters())
    n_params = sum(p.numel() for p in params)
    td = unroll_td(learner.policy, obs, act, rew_i, rew_e, beta, T)
    ((td.pow(2).mean(dim=0) * w).mean()).backward()
    full = [p.grad.detach().clone() for p in params]
    result = {"world": world, "params": n_params, "bytes": n_params * 4}

    modes = [("blocking", False, "nccl"), ("overlap", True, "nccl")]
    if cuda:                                            # the library's own NVLink all-reduce kernel (include/ngu_ar.h)
        modes += [("p2p_blocking", False, "p2p"), ("p2p_overlap", True, "p2p")]
    for key, overlap, transport in modes:
        learner, hy = build_learner()
        params = list(learner.policy.parameters())
        red = attach_r2d2_allreduce(learner, overlap=overlap, transport=transport)
        # the gradient the patched step produces, captured before the optimiser consumes it
        grads = {}
        orig = learner.optimizer.step
        learner.optimizer.step = lambda: grads.update(g=[p.grad.detach().clone() for p in params])
        for it in range(2):                             # twice: the second pass runs in the steady state (views in place)
            td = unroll_td(learner.policy, sl(obs), sl(act), sl(rew_i), sl(rew_e), beta[lo:hi], T)
            # equal slices: mean over the local slice, averaged over ranks == mean over the whole batch
            learner.step(td, w[lo:hi])
        learner.optimizer.step = orig
        # compare after the same clip the step applied (clip_grad_norm_ scales by a global factor <= 1)
        tot = torch.sqrt(sum((f.double() ** 2).sum() for f in full)).item()
        scale = min(1.0, hy["adam_clip_norm"] / (tot + 1e-6))
        worst_scaled = 0.0
        for a, b in zip(grads["g"], full):
            worst_scaled = max(worst_scaled, float((a - b * scale).abs().max() / (b.abs().max() * scale + 1e-12)))
        t = torch.tensor([worst_scaled], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ident = [p.grad.detach().clone() for p in params]
        same = True
        for x in ident:
            lo_, hi_ = x.clone(), x.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            same &= bool(torch.equal(lo_, hi_))
        result[key] = {"worst_rel_vs_full_batch": float(t.item()), "identical_across_ranks": same,
                       "copies_in_steady_state": red.copies_in, "buckets": len(red.buckets),
                       "update_count": learner.update_count, "transport": transport}
        assert t.item() < 2e-4, (key, t.item())
        assert same and red.copies_in == 0
        if transport == "p2p":
            assert red.p2p_timeouts() == 0
        del ident, grads
        red.close()

    torch.backends.cudnn.allow_tf32 = True
    # ---- timing: does the all-reduce hide behind backward? ----
    def timed(mode, iters=20 if cuda else 1, transport="nccl"):
        learner, hy = build_learner()
        params = list(learner.policy.parameters())
        red = GradAllReducer(params, bucket_bytes=(20 << 20) if mode == "overlap" else (64 << 20),
                             overlap=(mode == "overlap"), transport=transport) if mode != "none" else None
        Tt, Bl = 16, 32
        o = torch.randn(Tt, Bl, 1, 84, 84, device=dev)
        a = torch.randint(0, 18, (Tt, Bl), device=dev)
        r1 = torch.randn(Tt, Bl, 1, device=dev)
        r2 = torch.randn(Tt, Bl, 1, device=dev)
        bt = torch.zeros(Bl, 32, device=dev)
        ww = torch.ones(Bl, device=dev)
        tot = 0.0
        for it in range(iters + 3):
            td = unroll_td(learner.policy, o, a, r1, r2, bt, Tt)
            loss = (td.pow(2).mean(dim=0) * ww).mean()
            if red is not None:
                red.zero_grad()
            else:
                learner.optimizer.zero_grad(set_to_none=False)
            sync()
            dist.barrier()
            t0 = time.perf_counter()
            loss.backward()
            if red is not None:
                red()
            sync()
            if it >= 3:
                tot += time.perf_counter() - t0
        if red is not None:
            red.close()
        t = torch.tensor([tot / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    result["backward_ms"] = {"backward_only": round(timed("none"), 3), "backward_plus_blocking_allreduce": round(timed("blocking"), 3),
                             "backward_plus_overlapped_allreduce": round(timed("overlap"), 3),
                             **({"backward_plus_blocking_p2p_allreduce": round(timed("blocking", transport="p2p"), 3)} if cuda else {}),
                             "what": "reference DuelingLSTM, 16-step unroll, 32 sequences per rank; max over ranks"}

    # ---- raw bus bandwidth of the three buckets (gradients already in place: one ncclAllReduce(avg) each) ----
    bw = {}
    cases = [(name, n, "nccl") for name, n in {"embedding": 182866, "rnd_predictor": 473376, "r2d2": 8188595}.items()]
    if cuda:
        cases += [(name + "_p2p", n, "p2p") for name, n, _ in list(cases)]
    for name, n, transport in cases:
        ps = [torch.nn.Parameter(torch.zeros(n, device=dev))]
        red = GradAllReducer(ps, transport=transport)
        for _ in range(10):
            red()
        sync()
        dist.barrier()
        iters = 100 if cuda else 3
        t0 = time.perf_counter()
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for _ in range(iters):
            red()
        if cuda:
            e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / iters if cuda else (time.perf_counter() - t0) * 1e3 / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        bw[name] = {"bytes": n * 4, "ms": round(ms, 4), "algbw_GBs": round(n * 4 / ms / 1e6, 1),
                    "busbw_GBs": round(n * 4 / ms / 1e6 * 2 * (world - 1) / world, 1), "transport": transport}
        # correctness of this very buffer: rank-dependent values -> their mean, on every rank
        ps[0].grad.fill_(float(rank + 1))
        red()
        sync()
        assert torch.allclose(ps[0].grad, torch.full_like(ps[0].grad, (world + 1) / 2)), name
        red.close()
    result["allreduce"] = bw
    if rank == 0:
        print(json.dumps({"learner_dp_check": result}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
