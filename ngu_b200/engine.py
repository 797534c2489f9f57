"""EpisodicEngine — thin Python owner of one libngu_epi handle (one GPU, one shard of actors).

PyTorch is plumbing here: it owns the device buffers (so the episodic memory is
a normal ``torch.Tensor`` the caller can read, exactly like the reference's
``EpisodicNovelty.episodic_memory``) and provides the CUDA stream; all compute
is in the sm_100a kernels behind the C ABI (include/ngu_epi.h).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _cabi
from ._cabi import NguEpiCfg, NguEpiState, check

# model_hypr keys read by the reference (ngu/models/model_hypr.py:52-57)
DEFAULT_HYPR = dict(
    episodic_memory_capacity=30000,
    kernel_epsilon=0.0001,
    num_neighbors=10,
    cluster_distance=0.008,
    kernel_pseudo_counts_constant=0.001,
    kernel_maximum_similarity=8,
)


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class EpisodicEngine:
    """One actor shard on one B200.

    Args mirror ``EpisodicNovelty.__init__`` (episodic_novelty.py:9-35) after
    the hyper-parameter dict has been unpacked.  ``prior_count=None`` reproduces
    the reference's ``RunningMeanStd((k,))`` quirk (count starts at k).
    """

    def __init__(self, n_actors, capacity, k=10, kernel_epsilon=1e-4, cluster_distance=0.008,
                 pseudo_counts=1e-3, max_similarity=8.0, max_reward_scale=5.0, mode="reference",
                 device=None, prior_count=None, scan_variant=0, scan_ctas_per_sm=0):
        self._lib = _cabi.load()
        if not torch.cuda.is_available():
            raise _cabi.NguEpiError("ngu_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.n_actors, self.capacity, self.k = int(n_actors), int(capacity), int(k)
        self.mode = mode
        cfg = NguEpiCfg(
            n_actors=self.n_actors, capacity=self.capacity, dim=_cabi.DIM, k=self.k,
            mode={"reference": _cabi.MODE_REFERENCE, "paper": _cabi.MODE_PAPER}[mode],
            device=self.device.index,
            kernel_epsilon=kernel_epsilon, cluster_distance=cluster_distance, pseudo_counts=pseudo_counts,
            max_similarity=max_similarity, max_reward_scale=max_reward_scale, reserved0=0.0,
            prior_count=float(self.k if prior_count is None else prior_count),
            scan_ctas_per_sm=scan_ctas_per_sm, scan_variant=scan_variant)
        # torch owns the episodic memory: actor-major [B][N][32], zero-initialised
        self._memory = torch.zeros((self.n_actors, self.capacity, _cabi.DIM), dtype=torch.float32, device=self.device)
        h = C.c_void_p()
        rc = self._lib.ngu_epi_create(C.byref(cfg), _ptr(self._memory), C.byref(h))
        check(rc, None)
        self._h = h
        p = C.c_void_p()
        check(self._lib.ngu_epi_rank_sums(self._h, C.byref(p)), self._h)
        self.rank_sums = torch.as_tensor(_DevArray(p.value, (2 * self.k + 1,), "<f8"), device=self.device)
        self._r_int = torch.empty(self.n_actors, dtype=torch.float32, device=self.device)
        self._r_epi = torch.empty(self.n_actors, dtype=torch.float32, device=self.device)
        self.log_sums = torch.zeros(5, dtype=torch.float64, device=self.device)
        self._step_host_fn = self._lib.ngu_epi_step_host
        self._out = np.empty((2, self.n_actors), np.float32)      # step_host result staging (address cached)
        self._out_ptr = self._out.ctypes.data
        self._hp_cache = {}
        self._q_shape, self._a_shape = (self.n_actors, _cabi.DIM), (self.n_actors,)
        self._version = 0            # bumped by every call that can change the running statistics (state cache key)
        self._state_cache = (-1, None)

    # ---- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ngu_epi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- views ----------------------------------------------------------------
    @property
    def episodic_memory(self) -> torch.Tensor:
        """(capacity, n_actors, 32) strided view, the reference's layout (episodic_novelty.py:32-33)."""
        return self._memory.permute(1, 0, 2)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _q(self, q: torch.Tensor) -> torch.Tensor:
        if q.device != self.device or q.dtype != torch.float32 or not q.is_contiguous():
            q = q.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(q.shape) != (self.n_actors, _cabi.DIM):
            raise ValueError(f"controllable state must be ({self.n_actors}, {_cabi.DIM}), got {tuple(q.shape)}")
        return q

    # ---- device-pointer path (no host sync) -------------------------------------
    def scan(self, q: torch.Tensor):
        self._version += 1
        self._q_keep = self._q(q)
        check(self._lib.ngu_epi_scan(self._h, _ptr(self._q_keep), self._stream()), self._h)

    def finish(self, sums_global=None, alpha=None, want_log_sums=False):
        self._version += 1
        if alpha is not None:
            alpha = alpha.to(device=self.device, dtype=torch.float32).contiguous()
        self._alpha_keep = alpha
        check(self._lib.ngu_epi_finish(self._h, _ptr(sums_global), _ptr(alpha), _ptr(self._r_int), _ptr(self._r_epi),
                                       _ptr(self.log_sums) if want_log_sums else None, self._stream()), self._h)
        return self._r_int, self._r_epi

    def log_update(self, log_sums_global: torch.Tensor):
        self._version += 1
        check(self._lib.ngu_epi_log_update(self._h, _ptr(log_sums_global), self._stream()), self._h)

    def append(self, q: torch.Tensor):
        self._version += 1
        self._q_keep = self._q(q)
        check(self._lib.ngu_epi_append(self._h, _ptr(self._q_keep), self._stream()), self._h)

    def reset(self, done: torch.Tensor):
        done = done.to(device=self.device)
        if done.dtype == torch.bool:
            done = done.view(torch.uint8)            # same bytes (0 / 1): no cast kernel on the step's critical path
        elif done.dtype != torch.uint8:
            done = (done != 0).view(torch.uint8)
        done = done.contiguous()
        self._done_keep = done
        check(self._lib.ngu_epi_reset(self._h, _ptr(done), self._stream()), self._h)

    def step(self, q: torch.Tensor, alpha: torch.Tensor | None = None, pipelined: bool = False, out=None):
        """scan -> finish -> append on the current stream; returns device tensors (r_int, r_epi).

        ``pipelined=True`` (ngu_epi_step_pipelined): for steps enqueued back to back whose inputs are already on the
        device.  The step's scan overlaps the previous step's merge tail / epilogue / exchange; same results, bit
        for bit.  Contract (include/ngu_epi.h): nothing else was enqueued on the stream since the previous step,
        this step's q / alpha were ready before the previous step was enqueued, and the previous step's q stays
        intact until this step has completed (the engine keeps a reference to it for that long).
        ``out=(r_int, r_epi)``: write the rewards there instead of the engine's own pair of buffers."""
        self._version += 1
        qk = self._q(q)
        a_in = alpha
        if alpha is not None:
            alpha = alpha.to(device=self.device, dtype=torch.float32).contiguous()
        if pipelined and (qk is not q or (alpha is not None and alpha.data_ptr() != a_in.data_ptr())):
            # an input had to be converted: that copy is work enqueued between the two steps, and its result is not
            # "ready before the previous step" - the pipelined contract does not hold, launch the ordinary way
            pipelined = False
        r_int, r_epi = out if out is not None else (self._r_int, self._r_epi)
        if out is not None:
            for r in out:
                if r is not None and (r.device != self.device or r.dtype != torch.float32 or not r.is_contiguous()
                                      or r.numel() != self.n_actors):
                    raise ValueError("out tensors must be contiguous float32 device tensors of n_actors elements")
        fn = self._lib.ngu_epi_step_pipelined if pipelined else self._lib.ngu_epi_step
        check(fn(self._h, _ptr(qk), _ptr(alpha), _ptr(r_int), _ptr(r_epi), self._stream()), self._h)
        self._q_keep2, self._q_keep = getattr(self, "_q_keep", None), qk      # the previous q outlives this step
        self._alpha_keep2, self._alpha_keep = getattr(self, "_alpha_keep", None), alpha
        return r_int, r_epi

    def step_rnd(self, q: torch.Tensor, pred: torch.Tensor, targ: torch.Tensor):
        """ngu_epi_step_rnd: RND features in, the modulator is computed inside the epilogue kernel.
        Returns device tensors (r_int, r_epi, alpha)."""
        self._version += 1
        self._q_keep = self._q(q)
        pred = pred.to(device=self.device, dtype=torch.float32).contiguous()
        targ = targ.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(pred.shape) != (self.n_actors, 128) or tuple(targ.shape) != (self.n_actors, 128):
            raise ValueError("RND features must be (n_actors, 128)")
        self._rnd_keep = (pred, targ)
        if not hasattr(self, "_alpha_out"):
            self._alpha_out = torch.empty(self.n_actors, dtype=torch.float32, device=self.device)
        check(self._lib.ngu_epi_step_rnd(self._h, _ptr(self._q_keep), _ptr(pred), _ptr(targ), 128, _ptr(self._r_int),
                                         _ptr(self._r_epi), _ptr(self._alpha_out), self._stream()), self._h)
        return self._r_int, self._r_epi, self._alpha_out

    # ---- host-buffer path (what the drop-in calls with CPU tensors) ---------------
    def _host_ptr(self, x, shape, what):
        """Address of a host float32 C-contiguous array (CPU torch tensor or numpy array) of `shape`;
        anything else is converted first.  Returns (address, keep-alive object)."""
        if type(x) is torch.Tensor:
            if x.device.type != "cpu" or x.dtype != torch.float32 or not x.is_contiguous():
                x = x.detach().to(device="cpu", dtype=torch.float32).contiguous()
            if tuple(x.shape) != shape:
                raise ValueError(f"{what} must be {shape}, got {tuple(x.shape)}")
            return x.data_ptr(), x              # 0.1 us, where ndarray.ctypes.data costs 1.8 us
        if type(x) is not np.ndarray or x.dtype != np.float32 or not x.flags.c_contiguous:
            x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != shape:
            raise ValueError(f"{what} must be {shape}, got {x.shape}")
        return x.ctypes.data, x

    def _host_ptr_cached(self, x, shape, what):
        """_host_ptr with a small identity cache: a caller that cycles through the same few (pinned) buffers - the
        normal way to feed a synchronous call - pays the dtype / layout / shape checks once per buffer, not per step.
        A hit needs the very same live object (weak reference) still pointing at the same memory."""
        e = self._hp_cache.get(id(x))
        if e is not None and e[1]() is x and (e[2] or x.data_ptr() == e[0]):
            return e[0], x
        ptr, keep = self._host_ptr(x, shape, what)
        if keep is x:                                   # only buffers that were usable as they are
            if len(self._hp_cache) >= 64:
                self._hp_cache.clear()
            try:
                self._hp_cache[id(x)] = (ptr, weakref.ref(x), type(x) is not torch.Tensor)
            except TypeError:
                pass
        return ptr, keep

    def step_host(self, q, alpha=None, out=None):
        """ngu_epi_step_host: host float32 buffers in (CPU torch tensors or numpy arrays; torch tensors are
        the cheaper to pass), host float32 rewards (r_int, r_epi) out, synchronous.  The rewards are fresh numpy arrays
        unless the caller passes ``out``: a C-contiguous float32 numpy array of shape (2, n_actors) the library writes
        into directly (rows r_int, r_epi are returned as views of it).
        Called back to back, the library keeps the next step's launches pre-armed on the GPU (include/ngu_epi.h)."""
        q_ptr, _q = self._host_ptr_cached(q, self._q_shape, "controllable state")
        a_ptr, _a = (None, None) if alpha is None else self._host_ptr_cached(alpha, self._a_shape, "alpha")
        self._version += 1
        if out is not None:
            e = self._hp_cache.get(id(out))
            if e is None or e[1]() is not out:
                if type(out) is not np.ndarray or out.dtype != np.float32 or out.shape != (2, self.n_actors) or not out.flags.c_contiguous:
                    raise ValueError(f"out must be a C-contiguous float32 numpy array of shape (2, {self.n_actors})")
                e = (out.ctypes.data, weakref.ref(out), True)
                if len(self._hp_cache) >= 64:
                    self._hp_cache.clear()
                self._hp_cache[id(out)] = e
            o_ptr = e[0]
            rc = self._step_host_fn(self._h, q_ptr, a_ptr, o_ptr, o_ptr + 4 * self.n_actors)
            if rc:
                check(rc, self._h)
            return out[0], out[1]
        rc = self._step_host_fn(self._h, q_ptr, a_ptr, self._out_ptr, self._out_ptr + 4 * self.n_actors)
        if rc:
            check(rc, self._h)
        out = self._out.copy()
        return out[0], out[1]

    def reset_host(self, done: np.ndarray):
        d = np.ascontiguousarray(done, dtype=np.uint8)
        if d.shape != (self.n_actors,):
            raise ValueError(f"done must be ({self.n_actors},), got {d.shape}")
        check(self._lib.ngu_epi_reset_host(self._h, d.ctypes.data_as(C.c_void_p)), self._h)

    # ---- results / state ------------------------------------------------------------
    def topk(self):
        """Last scan's (d2 [B,k] f32, idx [B,k] i32), best first."""
        d2 = np.empty((self.n_actors, self.k), np.float32)
        idx = np.empty((self.n_actors, self.k), np.int32)
        check(self._lib.ngu_epi_topk_host(self._h, d2.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p)), self._h)
        return d2, idx

    def get_state(self, allow_failed=False) -> dict:
        """Running statistics + cursor (synchronous).  Raises if the sticky timeout counter is set
        (include/ngu_epi.h, "sticky error") unless ``allow_failed``."""
        s = NguEpiState()
        check(self._lib.ngu_epi_get_state(self._h, C.byref(s)), self._h)
        if s.exchange_timeouts != 0 and not allow_failed:
            raise _cabi.NguEpiError(
                f"{s.exchange_timeouts} k-NN merge / p2p exchange round(s) timed out: rewards since then are NaN and "
                "the replicated statistics may have diverged between ranks; recreate the engine")
        k = self.k
        return dict(ed_mean=np.array(s.ed_mean[:k]), ed_var=np.array(s.ed_var[:k]), ed_count=s.ed_count,
                    epi_mean=s.epi_mean, epi_var=s.epi_var, epi_count=s.epi_count,
                    intr_mean=s.intr_mean, intr_var=s.intr_var, intr_count=s.intr_count,
                    cursor=int(s.cursor), steps=int(s.steps), exchange_timeouts=int(s.exchange_timeouts),
                    ll_mean=s.ll_mean, ll_var=s.ll_var, ll_count=s.ll_count)

    def cached_state(self) -> dict:
        """get_state(), fetched once per change of the statistics: a logging step reads six tracker fields
        (each read used to cost two device synchronisations and a kernel launch)."""
        if self._state_cache[0] != self._version:
            self._state_cache = (self._version, self.get_state())
        return self._state_cache[1]

    def set_cursor(self, cursor: int):
        """Move only the ring cursor (EpisodicNovelty.memory_idx, episodic_novelty.py:34)."""
        self._version += 1
        check(self._lib.ngu_epi_set_cursor(self._h, int(cursor)), self._h)

    def set_state(self, st: dict):
        """Replace the given PUBLIC fields (read-modify-write; the library's own bookkeeping - pending log-only
        sums, exchange sequence numbers, the sticky timeout counter - is never touched)."""
        self._version += 1
        s = NguEpiState()
        check(self._lib.ngu_epi_get_state(self._h, C.byref(s)), self._h)
        for j in range(self.k):
            if "ed_mean" in st: s.ed_mean[j] = float(np.ravel(st["ed_mean"])[j])
            if "ed_var" in st: s.ed_var[j] = float(np.ravel(st["ed_var"])[j])
        for name in ("ed_count", "epi_mean", "epi_var", "epi_count", "intr_mean", "intr_var", "intr_count",
                     "ll_mean", "ll_var", "ll_count"):
            if name in st: setattr(s, name, float(st[name]))
        for name in ("cursor", "steps"):
            if name in st: setattr(s, name, int(st[name]))
        check(self._lib.ngu_epi_set_state(self._h, C.byref(s)), self._h)

    def load_memory(self, slot_major: np.ndarray):
        """Load a host array in the reference layout (capacity, n_actors, 32)."""
        m = np.ascontiguousarray(slot_major, dtype=np.float32)
        if m.shape != (self.capacity, self.n_actors, _cabi.DIM):
            raise ValueError(f"expected {(self.capacity, self.n_actors, _cabi.DIM)}, got {m.shape}")
        check(self._lib.ngu_epi_load_memory_host(self._h, m.ctypes.data_as(C.c_void_p)), self._h)

    def dump_memory(self) -> np.ndarray:
        out = np.empty((self.capacity, self.n_actors, _cabi.DIM), np.float32)
        check(self._lib.ngu_epi_dump_memory_host(self._h, out.ctypes.data_as(C.c_void_p)), self._h)
        return out

    # ---- multi-GPU p2p exchange ------------------------------------------------------------
    def p2p_export(self, world: int) -> bytes:
        buf = C.create_string_buffer(64)
        check(self._lib.ngu_epi_p2p_export(self._h, buf, int(world)), self._h)
        return buf.raw

    def p2p_connect(self, rank: int, world: int, handles: list):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        check(self._lib.ngu_epi_p2p_connect(self._h, int(rank), int(world), blob), self._h)

    def log_flush(self):
        """Collective (p2p mode): bring the one-step-late log-only trackers up to date."""
        self._version += 1
        check(self._lib.ngu_epi_log_flush(self._h, self._stream()), self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ngu_epi_launch_count(self._h))

    def scan_timing_begin(self, max_calls: int):
        check(self._lib.ngu_epi_scan_timing_begin(self._h, int(max_calls)), self._h)

    def scan_timing_end(self):
        """(total device ms of the timed scan-kernel launches, number of launches)."""
        ms, n = C.c_double(), C.c_int32()
        check(self._lib.ngu_epi_scan_timing_end(self._h, C.byref(ms), C.byref(n)), self._h)
        return ms.value, n.value

    def schedule_chunks(self):
        """Test aid: the scan kernel's work schedule as a list of (first tile, end tile) per chunk."""
        n = self.scan_geometry()["chunks"]
        out = (C.c_int32 * 2)()
        chunks = []
        for i in range(n):
            check(self._lib.ngu_epi_debug_chunk(self._h, i, out), self._h)
            chunks.append((out[0], out[1]))
        return chunks

    def scan_geometry(self) -> dict:
        g = (C.c_int32 * 8)()
        check(self._lib.ngu_epi_scan_geometry(self._h, g), self._h)
        return dict(grid=g[0], block=g[1], dyn_smem=g[2], static_tiles_per_warp=g[3], stages=g[4], tile_rows=g[5],
                    dynamic_chunk_tiles=g[6], chunks=g[7])
