"""Device-resident sequence collector (include/ngu_seq.h): the buffering half of ``NGUAgent.collect_sequence`` and
``NGUAgent.push_collected`` (reference: ngu/models/ngu/ngu_agent.py:56-94, :96-155) next to the intrinsic-reward path.

The reference keeps five ``(n_actors, L, ...)`` CPU tensors (``state``, ``prev_action``, ``curr_action``,
``intrinsic_reward``, ``extrinsic_reward``, ngu_agent.py:40-46) and per step writes the transition at every actor's
``local_mem_idx``, pushes the actors whose buffer is full to the replay memory with the augmented and n-step rewards,
carries the overlap and advances / resets the indices.  Here the same buffers, the index vector and the sequence
store live in HBM and one ``step`` call enqueues the four small kernels that do all of it; the rewards the episodic
engine produces stay on the device.  Attribute names follow the reference so a caller's code reads the same.
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import NguSeqCfg, NguSeqViews
from .engine import _DevArray


def _check(lib, rc, h):
    if rc != 0:
        msg = lib.ngu_seq_last_error(h)
        raise _cabi.NguEpiError(f"libngu_epi (ngu_seq) error {rc}: {msg.decode() if msg else '?'}")


class SequenceCollector:
    def __init__(self, n_actors, model_hypr, obs_shape=(1, 84, 84), explr_beta=None, discounts=None, device=None,
                 store_capacity=None):
        """``model_hypr`` keys as in ngu_agent.py:22-27: trace_length, burnin_period, n_step, overlapping_period,
        replay_capacity.  ``explr_beta`` / ``discounts``: the per-actor (n_actors, 1) tensors of ``R2D2Actor``
        (r2d2_actor.py:30-31); the powers ``discounts ** i`` are taken with torch here, as the reference does (:77)."""
        if not torch.cuda.is_available():
            raise _cabi.NguEpiError("SequenceCollector needs a CUDA device (B200); there is no CPU fallback")
        self._lib = _cabi.load()
        dev = torch.cuda.current_device() if device is None else (device if isinstance(device, int) else torch.device(device).index or 0)
        self.device = torch.device("cuda", dev)
        self.n_actors = int(n_actors)
        self.seq_len = int(model_hypr["trace_length"]) + int(model_hypr["burnin_period"])     # ngu_agent.py:24
        self.n_step = int(model_hypr["n_step"])
        self.mem_seq_len = self.seq_len + self.n_step                                            # :38
        self.overlap = int(model_hypr["overlapping_period"])
        self.obs_shape = tuple(obs_shape)
        self.obs_elems = int(np.prod(self.obs_shape))
        cap = int(store_capacity if store_capacity is not None else model_hypr.get("replay_capacity", 15000))
        self.capacity = max(cap, self.n_actors)
        cfg = NguSeqCfg(self.n_actors, self.seq_len, self.n_step, self.overlap, self.obs_elems, self.capacity, dev, 0)
        h = C.c_void_p()
        _check(self._lib, self._lib.ngu_seq_create(C.byref(cfg), C.byref(h)), None)
        self._h = h
        v = NguSeqViews()
        _check(self._lib, self._lib.ngu_seq_views_get(self._h, C.byref(v)), self._h)
        B, L, S, Cc = self.n_actors, self.mem_seq_len, self.seq_len, self.capacity
        view = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device=self.device)
        # the reference's attribute names (ngu_agent.py:36-46)
        self.state = view(v.state, (B, L) + self.obs_shape, "<f4")
        self.prev_action = view(v.prev_action, (B, L, 1), "<i8")
        self.curr_action = view(v.curr_action, (B, L, 1), "<i8")
        self.intrinsic_reward = view(v.intrinsic_reward, (B, L, 1), "<f4")
        self.extrinsic_reward = view(v.extrinsic_reward, (B, L, 1), "<f4")
        self.local_mem_idx = view(v.local_mem_idx, (B,), "<i8")
        # the sequence store: what PrioritizedReplayMemory.push copies (:135-138), ring of `capacity` sequences
        self.store = dict(
            state=view(v.st_state, (Cc, S) + self.obs_shape, "<f4"), prev_action=view(v.st_prev_action, (Cc, S, 1), "<i8"),
            curr_action=view(v.st_curr_action, (Cc, S, 1), "<i8"), intrinsic_reward=view(v.st_intrinsic_reward, (Cc, S, 1), "<f4"),
            extrinsic_reward=view(v.st_extrinsic_reward, (Cc, S, 1), "<f4"), augmented_reward=view(v.st_augmented_reward, (Cc, S, 1), "<f4"),
            nstep_reward=view(v.st_nstep_reward, (Cc, S, 1), "<f4"), actor=view(v.st_actor, (Cc,), "<i4"))
        self._counters = view(v.counters, (2,), "<i8")
        self._last_pushed = view(v.last_pushed, (1 + B,), "<i4")
        beta = torch.zeros(B, 1) if explr_beta is None else torch.as_tensor(explr_beta, dtype=torch.float32).reshape(B, 1).cpu()
        disc = torch.ones(B, 1) if discounts is None else torch.as_tensor(discounts, dtype=torch.float32).reshape(B, 1).cpu()
        pw = torch.cat([disc ** i for i in range(self.n_step)], dim=1).contiguous()               # r2d2_actor.py:77
        beta = beta.contiguous()
        _check(self._lib, self._lib.ngu_seq_set_factors(self._h, C.c_void_p(beta.data_ptr()), C.c_void_p(pw.data_ptr())), self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ngu_seq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, x, dtype, shape):
        x = torch.as_tensor(x)
        if x.device != self.device or x.dtype != dtype or not x.is_contiguous():
            x = x.to(device=self.device, dtype=dtype).contiguous()
        if x.numel() != int(np.prod(shape)):
            raise ValueError(f"expected {shape}, got {tuple(x.shape)}")
        return x

    def step(self, prev_obs, prev_act, action, intr_rew, extr_rew, done):
        """One environment step of the buffering (ngu_agent.py:76-82, push_collected, :93): tensors on any device
        (device tensors are used in place); enqueues on the current stream and returns nothing - read
        ``last_pushed()`` / ``store`` when the sequences are needed."""
        B = self.n_actors
        po = self._dev(prev_obs, torch.float32, (B,) + self.obs_shape)
        pa = self._dev(prev_act, torch.int64, (B, 1))
        ac = self._dev(action, torch.int64, (B, 1))
        ri = self._dev(intr_rew, torch.float32, (B, 1))
        re = self._dev(extr_rew, torch.float32, (B, 1))
        dn = torch.as_tensor(done)
        dn = (dn.to(self.device).view(torch.uint8) if dn.dtype == torch.bool else (dn.to(self.device) != 0).view(torch.uint8)).contiguous()
        self._keep = (po, pa, ac, ri, re, dn)
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(self._lib, self._lib.ngu_seq_step(self._h, p(po), p(pa), p(ac), p(ri), p(re), p(dn), stream), self._h)

    def last_pushed(self):
        """Store positions of the sequences the last step pushed (synchronises), in actor order."""
        lp = self._last_pushed.cpu().numpy()
        return lp[1:1 + int(lp[0])].copy()

    @property
    def pushed(self) -> int:
        return int(self._counters[0].item())

    def __len__(self):
        return min(self.pushed, self.capacity)
