"""torch-CPU port of the reference's hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

This is the CPU arm that bench.py times beside the GPU path ("cpu_baseline",
``--impl reference``).  The reference itself is Python over torch eager CPU ops
and cannot travel to the GPU box, so this file restates
``EpisodicNovelty.compute_episodic_novelty`` + the multiplier of
``IntrinsicNovelty.compute_intrinsic_novelty`` with the SAME sequence of ATen
calls the reference issues (ngu/models/intrinsic_novelty/episodic_novelty/
episodic_novelty.py:51-89, intrinsic_novelty.py:26-31), so that what is timed is
the reference's real CPU cost: broadcast sub -> pow -> sum -> sqrt -> topk on the
slot-major (N,B,32) memory, multi-threaded by ATen/OpenMP.  The running-mean
plumbing reuses the numpy restatement (oracle/episodic_oracle.py).

Pinned against the golden trajectories in tests/test_oracle.py.
"""
import numpy as np
import torch

from .episodic_oracle import RunningMeanStdOracle


class TorchCpuPort:
    def __init__(self, n_actors, capacity, k=10, kernel_epsilon=1e-4, cluster_distance=0.008,
                 pseudo_counts=0.001, max_similarity=8.0, max_rew=5.0):
        self.B, self.N, self.k = n_actors, capacity, k
        self.eps, self.cluster, self.c, self.s_max, self.max_rew = (
            kernel_epsilon, cluster_distance, pseudo_counts, max_similarity, max_rew)
        self.ed_rms = RunningMeanStdOracle(float(k), shape=(k,))           # episodic_novelty.py:28
        self.memory = torch.zeros((capacity, n_actors, 32))               # :32-33
        self.cursor = 0                                                   # :34

    @torch.no_grad()
    def step(self, q: torch.Tensor, alpha=None):
        B = self.B
        euc = ((self.memory - q.unsqueeze(0)) ** 2).sum(dim=-1).sqrt()    # :51-52
        kn = euc.topk(self.k, dim=0)                                      # :54
        self.ed_rms.update(kn.values.transpose(1, 0).numpy())             # :56
        mean = torch.from_numpy(self.ed_rms.mean).float()                 # pytorch_util.py:20
        nd = kn.values / mean.unsqueeze(-1)                               # :58
        cl = torch.max(nd - self.cluster, torch.zeros((self.k, B)))       # :60-65
        kv = self.eps / (cl + self.eps)                                   # :67-68
        sim = kv.sum(dim=0).sqrt() + self.c                               # :71-72
        r = 1 / sim                                                       # :75
        r[sim > self.s_max] = 0.0                                         # :76-78
        self.memory[self.cursor] = q.clone()                              # :88
        self.cursor = (self.cursor + 1) % self.N                          # :89
        if alpha is None:
            return r
        a = torch.as_tensor(alpha, dtype=torch.float64).clone()
        a[a < 1.0] = 1.0                                                  # intrinsic_novelty.py:26-27
        a[a > self.max_rew] = self.max_rew                                # :28-29
        return r * a                                                      # :31

    def reset(self, done):
        self.memory[:, torch.as_tensor(done, dtype=torch.bool), :] = 0.0  # intrinsic_novelty.py:39-40
