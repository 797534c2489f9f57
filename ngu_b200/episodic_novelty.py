"""Drop-in ``EpisodicNovelty`` (reference: ngu/models/intrinsic_novelty/episodic_novelty/
episodic_novelty.py:8-102) over the B200 episodic engine.

Same constructor, methods and externally read attributes as the reference
(SURVEY.md section 8b); the distance scan, top-k, running-mean update, reward
epilogue, append and reset run as sm_100a kernels behind the C ABI.  There is
no CPU path: using the module without a CUDA device raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import pytorch_util as ptu
from ._cabi import NguEpiError
from .embedding import Embedding
from .engine import EpisodicEngine
from .running_mean_std import DeviceRunningMeanStd


class EpisodicNovelty:
    def __init__(self, n_actors, n_act, obs_shape, model_hypr, logger, mode="reference"):
        self.n_actors, self.n_act, self.obs_shape = n_actors, n_act, obs_shape
        self.model_hypr, self.logger = model_hypr, logger
        self.controllable_state_dim = 32                       # episodic_novelty.py:23
        self.embedding = Embedding(n_act, obs_shape, self.controllable_state_dim, model_hypr, logger)
        self.capacity = model_hypr["episodic_memory_capacity"]
        self.mode = mode
        k = model_hypr["num_neighbors"]
        # device-resident trackers; ed_rms starts at count == k like RunningMeanStd((k,)) does upstream
        self.ed_rms = DeviceRunningMeanStd(self, "ed", True, float(k), k)
        self.epi_novel_rms = DeviceRunningMeanStd(self, "epi", False, 1e-4)
        self.update_count = 0
        self._engine = None
        self._memory_idx = 0
        self._pending_memory = None

    # ---- engine plumbing --------------------------------------------------------
    def _engine_or_none(self):
        return self._engine

    def _ensure_engine(self) -> EpisodicEngine:
        if self._engine is None:
            dev = ptu.current_device()
            if dev.type != "cuda":
                raise NguEpiError("EpisodicNovelty needs a CUDA device (B200); there is no CPU fallback")
            h = self.model_hypr
            self._engine = EpisodicEngine(
                self.n_actors, self.capacity, k=h["num_neighbors"], kernel_epsilon=h["kernel_epsilon"],
                cluster_distance=h["cluster_distance"], pseudo_counts=h["kernel_pseudo_counts_constant"],
                max_similarity=h["kernel_maximum_similarity"], mode=self.mode, device=dev.index or 0)
            if self._pending_memory is not None:
                self.episodic_memory = self._pending_memory
                self._pending_memory = None
        return self._engine

    @property
    def episodic_memory(self) -> torch.Tensor:
        """(capacity, n_actors, 32) view of the engine's HBM buffer (episodic_novelty.py:32-33)."""
        return self._ensure_engine().episodic_memory

    @episodic_memory.setter
    def episodic_memory(self, value: torch.Tensor):
        if self._engine is None and not torch.cuda.is_available():
            self._pending_memory = value
            return
        eng = self._ensure_engine()
        eng.episodic_memory.copy_(value.to(eng.device))
        # a torch copy is invisible to the library's stream bookkeeping: finish it before any *_host call can run
        torch.cuda.current_stream(eng.device).synchronize()

    @property
    def memory_idx(self) -> int:
        return self._memory_idx

    @memory_idx.setter
    def memory_idx(self, value: int):
        self._memory_idx = int(value) % self.capacity
        self._ensure_engine().set_cursor(self._memory_idx)

    # ---- hot path -------------------------------------------------------------------
    def _embed(self, obs) -> torch.Tensor:
        dev = self._ensure_engine().device
        return self.embedding(obs.to(dev)).detach()                      # episodic_novelty.py:45-47

    def _step(self, obs, alpha=None):
        """scan -> running mean -> reward (x clip(alpha)) -> append; device tensors out."""
        eng = self._ensure_engine()
        q = self._embed(obs)
        r_int, r_epi = eng.step(q, alpha)
        self._memory_idx = (self._memory_idx + 1) % self.capacity        # mirrors the device cursor
        return r_int, r_epi

    @torch.no_grad()
    def compute_episodic_novelty(self, obs):
        """Reference signature (episodic_novelty.py:37-84): FloatTensor[B] on the CPU."""
        _, r_epi = self._step(obs)
        return r_epi.cpu()

    def insert_state(self, controllable_state):
        """episodic_novelty.py:86-89."""
        self._ensure_engine().append(controllable_state)
        self._memory_idx = (self._memory_idx + 1) % self.capacity

    def knn(self):
        """(d2 [B,k], slot [B,k]) of the last query, best first (new contract; the
        reference discards topk's indices)."""
        return self._ensure_engine().topk()

    # ---- training / housekeeping (unchanged PyTorch) ------------------------------------
    def step(self, timestep_obs, timestep_next_obs, timestep_act):
        self.embedding.step(timestep_obs, timestep_next_obs, timestep_act)
        self.update_count += 1
        log = self.logger.log_scalar
        log("EpisodicNoveltyMean", self.epi_novel_rms.mean, self.update_count)
        log("EpisodicNoveltyVar", self.epi_novel_rms.var, self.update_count)
        log("KNeighborEucDistMean", np.mean(self.ed_rms.mean), self.update_count)
        log("KNeighborEucDistVar", np.mean(self.ed_rms.var), self.update_count)

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise NguEpiError("the episodic memory lives in B200 HBM; .to('cpu') is not supported")
        ptu.device = device if device.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.embedding.to(ptu.device)
        self._ensure_engine()
        return self

    # ---- checkpoint (the reference saves none of this) ------------------
    def state_dict(self):
        eng = self._ensure_engine()
        return {"engine": eng.get_state(), "memory": eng.episodic_memory.cpu().clone(),
                "embedding": self.embedding.state_dict()}

    def load_state_dict(self, sd):
        eng = self._ensure_engine()
        eng.episodic_memory.copy_(sd["memory"].to(eng.device))
        torch.cuda.current_stream(eng.device).synchronize()
        eng.set_state({k: v for k, v in sd["engine"].items() if k != "exchange_timeouts"})
        self._memory_idx = sd["engine"]["cursor"]
        self.embedding.load_state_dict(sd["embedding"])
