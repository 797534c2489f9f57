"""CUDA-graph capture of one whole intrinsic-novelty step.

The reference's step is ~40 tiny eager launches plus three host syncs.  Here
the complete device pipeline of ``IntrinsicNovelty.compute_intrinsic_novelty`` —

    H2D of the observations (pinned)  ->  Embedding CNN  ->  RND predictor/target CNNs
    ->  scan_topk kernel  ->  epilogue kernel (RND tail, running means, rewards, append)
    ->  D2H of the (B,) rewards (pinned)

— is captured once and replayed per step.  Replay is valid because everything that changes from
step to step lives in device memory (ring cursor, running statistics, the episodic memory itself)
and the observation / reward buffers are static.  The CNN weights are read at replay time, so
optimiser steps between replays are picked up.
"""
from __future__ import annotations

import torch


class GraphedIntrinsicStep:
    def __init__(self, intrinsic_novelty, obs_shape):
        self.inn = intrinsic_novelty
        eng = intrinsic_novelty.epi_novel._ensure_engine()
        self.eng = eng
        B = eng.n_actors
        self.obs_host = torch.empty((B, *obs_shape), dtype=torch.float32).pin_memory()
        self.obs_dev = torch.empty((B, *obs_shape), dtype=torch.float32, device=eng.device)
        self.r_host = torch.empty(B, dtype=torch.float32).pin_memory()
        self.graph = None
        self.replays = 0

    def _pipeline(self):
        inn, eng = self.inn, self.eng
        self.obs_dev.copy_(self.obs_host, non_blocking=True)
        pred, targ = inn.ll_novel(self.obs_dev)
        q = inn.epi_novel.embedding(self.obs_dev)
        r_int, _, _ = eng.step_rnd(q, pred, targ)
        self.r_host.copy_(r_int, non_blocking=True)

    def capture(self):
        """Capture WITHOUT side effects on the episodic state: the warm-up / capture passes run on a
        snapshot that is restored afterwards (capture itself executes nothing)."""
        eng = self.eng
        state = eng.get_state()
        mem = eng._memory.clone()
        self.obs_host.zero_()
        side = torch.cuda.Stream(device=eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):                      # cuDNN autotune / lazy init outside the capture
                self._pipeline()
        torch.cuda.current_stream(eng.device).wait_stream(side)
        torch.cuda.synchronize(eng.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._pipeline()
        torch.cuda.synchronize(eng.device)
        eng._memory.copy_(mem)
        eng.set_state({k: state[k] for k in state if k != "exchange_timeouts"})
        torch.cuda.synchronize(eng.device)
        return self

    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        """obs: CPU float tensor (B, *obs_shape) -> (B,1) CPU rewards (a fresh tensor)."""
        if self.graph is None:
            self.capture()
        self.obs_host.copy_(obs)
        self.graph.replay()
        torch.cuda.current_stream(self.eng.device).synchronize()
        self.replays += 1
        return self.r_host.clone().unsqueeze(-1)
