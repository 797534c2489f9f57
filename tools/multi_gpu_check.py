"""Multi-rank parity check: actor shards + rank-sum exchange vs the unsharded oracle.

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py
  (NCCL, one rank per GPU), or spawned by tests/test_gpu_sharded.py with --backend gloo --same-gpu
  (every rank on cuda:0, exchange over gloo) when only one GPU is present.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--actors", type=int, default=13)
    ap.add_argument("--capacity", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--exchange", default="auto")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(args.backend)
    from ngu_b200.sharding import ShardedEpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle

    B, N = args.actors, args.capacity
    rng = np.random.default_rng(2024)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[rng.random(N) < 0.2] = 0.0
    sh = ShardedEpisodicEngine(B, N, device=local, exchange="collective" if args.same_gpu else args.exchange)
    sh.engine.load_memory(mem0[:, sh.lo:sh.hi])
    orc = EpisodicOracle(B, N, use_c_scan=True)
    orc.memory[...] = mem0
    worst = 0.0
    for t in range(args.steps):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)
        done = rng.random(B) < 0.2
        want = orc.step(q, alpha)
        r_int, _ = sh.step(torch.from_numpy(q[sh.lo:sh.hi]).cuda(), torch.from_numpy(alpha[sh.lo:sh.hi]).cuda())
        r = r_int.cpu().numpy()
        d2, idx = sh.engine.topk()
        assert np.array_equal(idx.T, want["idx"][:, sh.lo:sh.hi]), f"rank {rank} step {t}: kNN slots differ"
        assert np.array_equal(d2.T.view(np.uint32), want["d2"][:, sh.lo:sh.hi].view(np.uint32))
        rel = np.abs(r - want["reward"][sh.lo:sh.hi]) / np.abs(want["reward"][sh.lo:sh.hi])
        worst = max(worst, float(rel.max()))
        assert rel.max() <= 1e-5, f"rank {rank} step {t}: reward rel err {rel.max():.2e}"
        orc.reset(done)
        sh.reset(torch.from_numpy(done[sh.lo:sh.hi]))
    n_steps = args.steps
    if sh.exchange == "p2p":     # bursts of software-pipelined steps: the exchange of step t runs under the scan of step t+1
        for steps in (2, 9, 30):
            qh = rng.standard_normal((steps, B, 32), dtype=np.float32)
            ah = (1.0 + 2.0 * rng.standard_normal((steps, B))).astype(np.float32)
            qd, ad = torch.from_numpy(qh[:, sh.lo:sh.hi].copy()).cuda(), torch.from_numpy(ah[:, sh.lo:sh.hi].copy()).cuda()
            r = torch.empty((steps, sh.hi - sh.lo), dtype=torch.float32, device="cuda")
            e = torch.empty_like(r)
            torch.cuda.synchronize()
            for t in range(steps):
                sh.step(qd[t], ad[t], pipelined=True, out=(r[t], e[t]))
            torch.cuda.synchronize()
            got = r.cpu().numpy()
            for t in range(steps):
                want = orc.step(qh[t], ah[t])
                rel = np.abs(got[t] - want["reward"][sh.lo:sh.hi]) / np.abs(want["reward"][sh.lo:sh.hi])
                worst = max(worst, float(rel.max()))
                assert rel.max() <= 1e-5, f"rank {rank} pipelined burst of {steps}, step {t}: reward rel err {rel.max():.2e}"
            d2, idx = sh.engine.topk()
            assert np.array_equal(idx.T, want["idx"][:, sh.lo:sh.hi]), f"rank {rank}: kNN slots differ after a pipelined burst"
            assert np.array_equal(d2.T.view(np.uint32), want["d2"][:, sh.lo:sh.hi].view(np.uint32))
            n_steps += steps
    st = sh.get_state()          # collective in p2p mode: flushes the one-step-late log-only sums
    np.testing.assert_allclose(st["ed_mean"], orc.ed_rms.mean, rtol=1e-6)
    if sh.exchange != "collective" or sh.track_log_stats:
        # log-only trackers saw every actor of every step (global count), on every rank
        assert abs(st["epi_count"] - (1e-4 + n_steps * B)) < 1e-6, st["epi_count"]
        assert abs(st["intr_count"] - (1e-4 + n_steps * B)) < 1e-6
    assert st["ed_count"] == float(orc.ed_rms.count)            # global actor count went into the tracker
    assert np.array_equal(sh.engine.dump_memory(), orc.memory[:, sh.lo:sh.hi])
    # every rank holds the same replicated statistics
    m = torch.from_numpy(st["ed_mean"]).cuda() if args.backend == "nccl" else torch.from_numpy(st["ed_mean"])
    lo, hi = m.clone(), m.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "running mean diverged between ranks"
    assert st["exchange_timeouts"] == 0
    print(f"rank {rank}/{world}: sharded parity ok ({sh.exchange}), actors [{sh.lo},{sh.hi}), worst reward rel err {worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
