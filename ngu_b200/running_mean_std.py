"""Running mean / variance trackers with the reference's interface.

``RunningMeanStd`` is the host-side mirror of ngu/utils/mpi_util.py:36-73 (same
constructor, ``mean`` / ``var`` / ``count`` attributes, ``update`` /
``update_from_moments``).  The reference reduces batch moments over
``MPI.COMM_WORLD``; here the optional ``comm`` is a ``torch.distributed``
process group (or None for a single process), and the reduction is one
all-reduce of ``[sum, sum of squares, count]`` instead of two MPI rounds.

``DeviceRunningMeanStd`` is a read-only proxy over the statistics that live on
the GPU inside the episodic engine (``ed_rms``, ``epi_novel_rms``,
``intrinsic_novel_rms``): they are updated by the finish kernel and only
fetched when somebody logs them.
"""
from __future__ import annotations

import numpy as np


def _allreduce_sum(vec: np.ndarray, comm) -> np.ndarray:
    import torch
    import torch.distributed as dist
    if comm is None and not (dist.is_available() and dist.is_initialized()):
        return vec
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    backend = dist.get_backend(comm)
    if backend == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=comm)
    return t.cpu().numpy()


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=(), comm=None, use_mpi=True):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        # The reference stores whatever arrives in the `epsilon` slot, so
        # RunningMeanStd((k,)) starts with count == k.
        self.count = float(np.ravel(np.asarray(epsilon, dtype=np.float64))[0])
        self.use_mpi = use_mpi
        self.comm = comm

    def update(self, x):
        x = np.asarray(x)
        if x.ndim == 0:                       # scalar sample (losses): mean=x, "std"=x as upstream does
            self.update_from_moments(float(x), float(x) ** 2, 1)
            return
        n_local = x.shape[0]
        if not self.use_mpi:
            # mpi_util.py:51-52: np.mean / np.std in the INPUT dtype (float32 batches give float32 moments),
            # then np.square(batch_std) (:56)
            self.update_from_moments(np.mean(x, axis=0), np.square(np.std(x, axis=0)), n_local)
            return
        x64 = x.astype(np.float64).reshape(n_local, -1)
        if self.use_mpi:
            width = x64.shape[1]
            packed = np.concatenate([x64.sum(axis=0), np.square(x64).sum(axis=0), [float(n_local)]])
            packed = _allreduce_sum(packed, self.comm)
            n = packed[-1]
            mean = packed[:width] / n
            var = np.maximum(packed[width:2 * width] / n - np.square(mean), 0.0)
        else:
            n = float(n_local)
            mean, var = x64.mean(axis=0), x64.var(axis=0)
        shape = x.shape[1:]
        self.update_from_moments(mean.reshape(shape), var.reshape(shape), n)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        """Chan et al. parallel merge (mpi_util.py:59-73)."""
        delta = batch_mean - self.mean
        total = self.count + batch_count
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / total
        self.mean = self.mean + delta * batch_count / total
        self.var = m2 / total
        self.count = total


class DeviceRunningMeanStd:
    """mean / var / count of one of the engine's device-resident trackers."""

    def __init__(self, owner, prefix: str, vector: bool, prior_count: float, k: int = 1):
        self._owner, self._prefix, self._vector = owner, prefix, vector
        self._prior, self._k = prior_count, k

    def _get(self, field):
        eng = self._owner._engine_or_none()
        if eng is None:       # nothing has run yet: priors (mpi_util.py:39-42)
            if field == "count":
                return self._prior
            base = np.zeros(self._k) if field == "mean" else np.ones(self._k)
            return base if self._vector else float(base[0])
        return eng.cached_state()[f"{self._prefix}_{field}"]

    mean = property(lambda self: self._get("mean"))
    var = property(lambda self: self._get("var"))
    count = property(lambda self: self._get("count"))
