"""world_size-2 gloo test (CPU) of the multi-GPU host logic: actor partition + rank-sum exchange.

The per-rank compute is stood in for by an oracle-backed engine with the same
interface as ngu_b200.engine.EpisodicEngine (tests may use the oracle); what is
under test is ngu_b200.sharding: that sharded steps with the exchange reproduce
the unsharded run (indices bit-exact, rewards <= 1e-5 relative).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleBackedEngine:
    """CPU stand-in with EpisodicEngine's scan/finish/append/reset interface."""

    def __init__(self, n_actors, capacity, device=None, k=10, **_):
        from oracle.episodic_oracle import EpisodicOracle
        self.o = EpisodicOracle(n_actors, capacity, k=k)
        self.k, self.n_actors, self.device = k, n_actors, torch.device("cpu")
        self.rank_sums = torch.zeros(2 * k + 1, dtype=torch.float64)
        self.log_sums = torch.zeros(5, dtype=torch.float64)

    def scan(self, q):
        from oracle.episodic_oracle import pairwise_sum_f32
        self.idx, self.d2 = self.o.scan(q.numpy())
        v = np.sqrt(self.d2).T                                  # (B,k)
        self.v = v
        self.rank_sums[:self.k] = torch.from_numpy(pairwise_sum_f32(v).astype(np.float64))
        self.rank_sums[self.k:2 * self.k] = torch.from_numpy(np.square(v.astype(np.float64)).sum(0))
        self.rank_sums[2 * self.k] = self.n_actors

    def finish(self, sums_global, alpha=None, want_log_sums=False):
        s = sums_global.numpy()
        k = self.k
        cnt = np.float32(s[2 * k])
        b_mean = (s[:k].astype(np.float32) / cnt)
        msd = np.maximum(s[k:2 * k] / s[2 * k] - 2 * b_mean.astype(np.float64) * s[:k] / s[2 * k] + b_mean.astype(np.float64) ** 2, 0)
        b_var = np.square(np.sqrt(msd.astype(np.float32)))
        self.o.ed_rms.update_from_moments(b_mean, b_var, cnt)
        o = self.o
        mean32 = o.ed_rms.mean.astype(np.float32)
        nd = self.v.T / mean32[:, None]
        cl = np.maximum(nd - o.cluster, np.float32(0))
        kv = (np.float32(1) / (cl + o.eps)) * o.eps
        acc = kv[0].copy()
        for j in range(1, k):
            acc = acc + kv[j]
        sim = np.sqrt(acc) + o.c
        r = np.where(sim > o.s_max, np.float32(0), np.float32(1) / sim).astype(np.float32)
        a = np.ones_like(r) if alpha is None else np.clip(alpha.numpy().astype(np.float32), 1.0, 5.0)
        return torch.from_numpy(r * a), torch.from_numpy(r)

    def log_update(self, log):
        pass

    def append(self, q):
        self.o.append(q.numpy())

    def step(self, q, alpha=None):
        self.scan(q)
        out = self.finish(self.rank_sums, alpha)
        self.append(q)
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngu_b200.sharding import ShardedEpisodicEngine
    sh = ShardedEpisodicEngine(B, N, engine_factory=OracleBackedEngine, track_log_stats=False)
    rng = np.random.default_rng(42)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    sh.engine.o.memory[...] = mem0[:, sh.lo:sh.hi]
    rewards, idxs = [], []
    for t in range(steps):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1 + 2 * rng.standard_normal(B)).astype(np.float32)
        r_int, _ = sh.step(torch.from_numpy(q[sh.lo:sh.hi]), torch.from_numpy(alpha[sh.lo:sh.hi]))
        rewards.append(r_int.numpy().copy())
        idxs.append(sh.engine.idx.copy())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r=np.stack(rewards), idx=np.stack(idxs), lo=sh.lo, hi=sh.hi)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_exchange_equals_unsharded(tmp_path):
    from oracle.episodic_oracle import EpisodicOracle
    B, N, steps, world = 7, 120, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, N, steps, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    ref = EpisodicOracle(B, N)
    ref.memory[...] = rng.standard_normal((N, B, 32), dtype=np.float32)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [0, 4] and [int(p["hi"]) for p in parts] == [4, 7]
    for t in range(steps):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1 + 2 * rng.standard_normal(B)).astype(np.float32)
        want = ref.step(q, alpha)
        got_r = np.concatenate([p["r"][t] for p in parts])
        got_idx = np.concatenate([p["idx"][t] for p in parts], axis=1)
        assert np.array_equal(got_idx, want["idx"])
        rel = np.abs(got_r - want["reward"]) / np.abs(want["reward"])
        assert rel.max() <= 1e-5, rel.max()
