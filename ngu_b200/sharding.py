"""Actor shards over the GPUs of one box + the per-step rank-sum exchange.

The scan, top-k, append and reset are independent per actor (each actor only
touches its own memory column: reference episodic_novelty.py:51-52,88 and
intrinsic_novelty.py:39), so actors are block-partitioned over ranks, one
process per GPU.  The only coupling is the batch mean of the k-NN distances
over ALL actors that feeds ``ed_rms`` (episodic_novelty.py:56 -> mpi_util.py:7-18):
2k+1 doubles per step, exchanged between the scan and the finish kernel.  This
is exactly what the reference's ``MPI.Allreduce`` does (mpi_util.py:17), here as
an exchange over NVLink peer memory FUSED INTO THE EPILOGUE KERNEL (exchange="p2p":
LL-tagged peer stores, csrc/p2p_exchange.cuh; the log-only sums of
episodic_novelty.py:82 / intrinsic_novelty.py:32 ride on the next step's round), or,
as the fallback transport (exchange="collective": gloo tests, no peer access), as
one torch.distributed all-reduce between a split scan and finish plus a second tiny
one for the log-only statistics.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .engine import EpisodicEngine


def actor_partition(n_actors: int, world_size: int):
    """Contiguous block partition: rank r owns [bounds[r], bounds[r+1])."""
    base, rem = divmod(n_actors, world_size)
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def exchange_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a small tensor over the ranks of `group` (in place); no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardedEpisodicEngine:
    """This rank's shard of a `global_actors`-actor episodic memory."""

    def __init__(self, global_actors, capacity, group=None, device=None, track_log_stats=True,
                 engine_factory=EpisodicEngine, exchange="auto", **engine_kwargs):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = actor_partition(global_actors, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if self.hi == self.lo:
            raise ValueError("more ranks than actors: every rank needs at least one actor")
        # engine_factory is a seam for the CPU (gloo) tests of this exchange logic; the product
        # always runs the CUDA engine
        self.engine = engine_factory(self.hi - self.lo, capacity, device=device, **engine_kwargs)
        self.track_log_stats = track_log_stats
        k = self.engine.k
        self._sums = torch.zeros(2 * k + 1, dtype=torch.float64, device=self.engine.device)
        self._log = torch.zeros(5, dtype=torch.float64, device=self.engine.device)
        # exchange transport: "p2p" = NVLink peer stores fused into the finish kernel (one process per
        # GPU, distinct GPUs), "collective" = torch.distributed all-reduce between scan and finish
        if exchange == "auto":
            exchange = "p2p" if (self.world > 1 and dist.get_backend(group) == "nccl"
                                 and hasattr(self.engine, "p2p_export")) else "collective"
        self.exchange = exchange if self.world > 1 else "none"
        if self.exchange == "off":          # timing experiments only: ranks run independently (wrong global mean)
            self.world = 1
        if self.exchange == "p2p":
            mine = self.engine.p2p_export(self.world)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.engine.p2p_connect(self.rank, self.world, handles)
            dist.barrier(group=group)

    def local_slice(self, x):
        return x[self.lo:self.hi]

    def step(self, q_local: torch.Tensor, alpha_local: torch.Tensor | None = None, pipelined: bool = False, out=None):
        """One hot-path step for this shard; returns device tensors (r_int, r_epi) of the local actors.
        ``pipelined``: EpisodicEngine.step's software-pipelined launch (back-to-back steps with device-resident
        inputs; with the p2p exchange the previous step's NVLink rendezvous then runs under this step's scan)."""
        eng = self.engine
        if self.world == 1 or self.exchange == "p2p":
            if pipelined or out is not None:
                return eng.step(q_local, alpha_local, pipelined=pipelined, out=out)
            return eng.step(q_local, alpha_local)      # p2p: the exchange happens inside the epilogue kernel
        eng.scan(q_local)
        self._sums.copy_(eng.rank_sums)
        exchange_sum(self._sums, self.group)                       # mpi_util.py:17
        out = eng.finish(self._sums, alpha_local, want_log_sums=self.track_log_stats)
        if self.track_log_stats:
            self._log.copy_(eng.log_sums)
            exchange_sum(self._log, self.group)
            eng.log_update(self._log)
        eng.append(q_local)
        return out

    def reset(self, done_local: torch.Tensor):
        self.engine.reset(done_local)

    def get_state(self) -> dict:
        """Collective in p2p mode (flushes the deferred log-only sums first)."""
        if self.exchange == "p2p":
            self.engine.log_flush()
        return self.engine.get_state()
