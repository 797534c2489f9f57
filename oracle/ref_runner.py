"""Drive the UNMODIFIED reference (staged in oracle/_ref by oracle/build.py) — BASELINE INFRASTRUCTURE ONLY.

What bench.py times as the reference arm: the imported ``ngu.models.intrinsic_novelty.IntrinsicNovelty``
(ngu/models/intrinsic_novelty/intrinsic_novelty.py:9-55) exactly as SURVEY.md section 8c/8d prescribes:

* synthetic controllable states: ``epi_novel.embedding = identity`` so that ``obs`` is the (B,32) state
  (``len(obs)``, ``.to``, ``.detach()`` all work, episodic_novelty.py:44-47) and a known RND modulator
  injected through ``ll_novel.compute_lifelong_curiosity`` — the same two seams
  tests/golden/make_golden.py uses;
* or the real CNN producers (config-1 substitute: observations in, nothing replaced);
* ``device='cpu'`` is the reference's CPU path (the headline baseline); ``device='cuda'`` runs the very
  same module with ``ptu.device`` on the GPU — the torch-eager ATen kernels minoring/NGU launches when it
  is given a GPU, "the kernel to beat on the same box" (SURVEY.md section 8d).

Nothing under ngu_b200/ imports this module.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ngu", "models", "intrinsic_novelty", "intrinsic_novelty.py"))


def import_reference(use_gpu=False):
    """Import the staged reference; returns (torch, IntrinsicNovelty, model_hypr, ptu)."""
    if not available():
        raise ImportError("oracle/_ref is not staged: run __graft_entry__.build() where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)          # ahead of everything: provides `ngu` and the mpi4py stand-in
    import torch
    import ngu.utils.pytorch_util as ptu
    ptu.init_device(use_gpu=use_gpu)
    from ngu.models.intrinsic_novelty import IntrinsicNovelty
    from ngu.models.model_hypr import model_hypr
    return torch, IntrinsicNovelty, model_hypr, ptu


class _NullLogger:
    def log_scalar(self, *a, **k):
        pass


class ReferenceRunner:
    """One reference ``IntrinsicNovelty`` on `device`, fed synthetic controllable states (identity
    embedding + injected alpha) or real observations (``real_cnn=True``)."""

    def __init__(self, n_actors, capacity, device="cpu", real_cnn=False, threads=None, seed=1234):
        torch, IntrinsicNovelty, model_hypr, ptu = import_reference(use_gpu=(device == "cuda"))
        self.torch, self.ptu = torch, ptu
        if threads:
            torch.set_num_threads(int(threads))
        self.B, self.N = int(n_actors), int(capacity)
        hy = dict(model_hypr, episodic_memory_capacity=self.N)
        torch.manual_seed(seed)
        self.inn = IntrinsicNovelty(self.B, 18, (1, 84, 84), hy, _NullLogger())
        self.inn.to(ptu.device)
        self.real_cnn = real_cnn
        if not real_cnn:
            self.inn.epi_novel.embedding = lambda x: x
        self.k = hy["num_neighbors"]

    def fill_memory(self, seed=1234):
        """Fully populated N(0,1) memory, reference layout (N,B,32)."""
        torch = self.torch
        g = torch.Generator().manual_seed(seed)
        mem = torch.randn(self.N, self.B, 32, generator=g)
        self.inn.epi_novel.episodic_memory = mem.to(self.ptu.device)

    def set_memory(self, mem_slot_major: np.ndarray):
        self.inn.epi_novel.episodic_memory = self.torch.from_numpy(np.ascontiguousarray(mem_slot_major)).to(self.ptu.device)

    def step(self, obs, alpha=None):
        """THE reference call: compute_intrinsic_novelty(obs) -> (B,1) CPU tensor."""
        torch = self.torch
        if not self.real_cnn:
            a = np.ones(self.B) if alpha is None else np.asarray(alpha, dtype=np.float64)
            self.inn.ll_novel.compute_lifelong_curiosity = lambda o, a=a: torch.from_numpy(a.copy())
        return self.inn.compute_intrinsic_novelty(obs)

    def reset(self, done):
        self.inn.reset_memory_if_done(self.torch.as_tensor(np.asarray(done, dtype=bool)))


def cpu_model_string() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    import platform
    return platform.processor() or "unknown"
