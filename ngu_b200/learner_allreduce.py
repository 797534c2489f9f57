"""Learner gradient all-reduce hook (BASELINE north_star "the learner's
embedding/RND/R2D2 gradient allreduce runs over NCCL on NVLink").

The reference has NO gradient averaging anywhere: this is new, data-parallel
functionality attached at the three seams where gradients exist and clipping has not happened yet:

    ngu/models/intrinsic_novelty/episodic_novelty/embedding.py:60-61        (182,866 params,  0.73 MB)
    ngu/models/intrinsic_novelty/lifelong_novelty/lifelong_novelty.py:64-65 (473,376 params,  1.89 MB)
    ngu/models/r2d2/r2d2_learner.py:53-54                                   (8,188,595 params, 32.75 MB)

``GradAllReducer`` owns a few flat, contiguous buckets and makes every parameter's ``.grad`` a VIEW
of its bucket, so autograd accumulates straight into the communication buffer: a reduction is one
``ncclAllReduce(avg)`` per bucket, with no copy in, no copy out and no divide kernel (round 1 copied
both ways and divided: 281 GB/s bus bandwidth on the 32 MB bucket where the fabric gives > 700).
With ``overlap=True`` a bucket is reduced on a side stream as soon as autograd has produced all of
its gradients (post-accumulate-grad hooks, last layers first), so the transfer of the R2D2 bucket
hides behind the rest of backward; the call at the seam then only waits.

At a seam (what ``Embedding.step`` / ``LifelongNovelty.step`` of this package do through their
``grad_hook`` attribute, and what ``attach_r2d2_allreduce`` installs on an ``R2D2Learner``):

    reducer = GradAllReducer(params)            # once
    reducer.zero_grad()                          # instead of optimizer.zero_grad(): keeps the views
    loss.backward()
    reducer(params)                              # <- new line: average over ranks
    nn.utils.clip_grad_norm_(params, 40)
    optimizer.step()

``optimizer.zero_grad()`` (which drops ``.grad`` to None by default) still works: the reducer then
copies the fresh gradient tensors into the bucket once and re-installs the views.
"""
from __future__ import annotations

import types

import torch
import torch.distributed as dist


class _P2PBucket:
    """One flat bucket in a buffer libngu_epi allocates and the peer ranks map over CUDA IPC (include/ngu_ar.h)."""

    def __init__(self, numel, device, group):
        import ctypes as C
        from . import _cabi
        from .engine import _DevArray
        self._C, self._lib = C, _cabi.load()
        h = C.c_void_p()
        self._check(self._lib.ngu_ar_create(int(numel), device.index, C.byref(h)), None)
        self._h = h
        ptr = C.c_void_p()
        self._check(self._lib.ngu_ar_buffer(self._h, C.byref(ptr)), self._h)
        self.flat = torch.as_tensor(_DevArray(ptr.value, (int(numel),), "<f4"), device=device)
        mine = C.create_string_buffer(_cabi.AR_HANDLE_BYTES)
        self._check(self._lib.ngu_ar_export(self._h, mine), self._h)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handles = [None] * world
        dist.all_gather_object(handles, mine.raw, group=group)
        self._check(self._lib.ngu_ar_connect(self._h, rank, world, b"".join(handles)), self._h)
        dist.barrier(group=group)

    def _check(self, rc, h):
        if rc != 0:
            msg = self._lib.ngu_ar_last_error(h)
            raise RuntimeError(f"libngu_epi (ngu_ar) error {rc}: {msg.decode() if msg else '?'}")

    def allreduce_avg(self, stream):
        self._check(self._lib.ngu_ar_allreduce_avg(self._h, self._C.c_void_p(stream.cuda_stream)), self._h)

    def timeouts(self):
        return int(self._lib.ngu_ar_timeouts(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.flat = None
            self._lib.ngu_ar_destroy(self._h)
            self._h = None


class GradAllReducer:
    def __init__(self, params, group=None, bucket_bytes=64 << 20, overlap=False, transport="nccl"):
        """``transport``: "nccl" (default; north_star: NCCL over NVLink - or whatever backend the process group has) or
        "p2p": the library's own all-reduce kernel over NVLink peer memory (include/ngu_ar.h; one process per GPU,
        at most 8 ranks on one box, float32 CUDA parameters): one pass, each element crosses NVLink once in and once
        out, result bit-identical on all ranks."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.transport = transport if self.world > 1 else "nccl"
        if self.transport not in ("nccl", "p2p"):
            raise ValueError("transport must be 'nccl' or 'p2p'")
        if self.transport == "p2p" and (self.world > 8 or any(p.dtype != torch.float32 or p.device.type != "cuda" for p in self.params)):
            raise ValueError("transport='p2p' needs float32 CUDA parameters and at most 8 ranks")
        self._p2p = []
        self._avg_native = (dist.is_initialized() and dist.get_backend(group) == "nccl")   # gloo has no AVG
        # buckets in reverse order: autograd finishes the last layers first
        self.buckets = []          # list of (flat tensor, [params], [views])
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._pending = []
        self._ready = {}
        self._hooks = []
        self._stream = None
        self.copies_in = 0          # gradients that had to be copied into their bucket (0 in the steady state)
        if self.overlap:
            dev = self.params[0].device
            self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
            for bi, (_, ps, _) in enumerate(self.buckets):
                for p in ps:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _close(self, ps):
        numel = sum(p.numel() for p in ps)
        if self.transport == "p2p":
            pb = _P2PBucket(numel, ps[0].device, self.group)
            self._p2p.append(pb)
            flat = pb.flat
        else:
            self._p2p.append(None)
            flat = torch.zeros(numel, dtype=ps[0].dtype, device=ps[0].device)
        views, off = [], 0
        for p in ps:
            v = flat[off:off + p.numel()].view_as(p)
            views.append(v)
            off += p.numel()
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v                       # autograd accumulates in place from now on
        self.buckets.append((flat, ps, views))

    @property
    def nbytes(self):
        return sum(f.numel() * f.element_size() for f, _, _ in self.buckets)

    def zero_grad(self):
        """Zero the buckets and keep every ``.grad`` a view of its bucket (use instead of optimizer.zero_grad())."""
        for flat, ps, views in self.buckets:
            flat.zero_()
            for p, v in zip(ps, views):
                if p.grad is not v:
                    p.grad = v

    def _adopt(self, bi):
        """Gradients that are not views of the bucket (the caller dropped them with set_to_none): copy them in once
        and re-install the views."""
        flat, ps, views = self.buckets[bi]
        src, dst = [], []
        for p, v in zip(ps, views):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
                p.grad = v
        if src:
            torch._foreach_copy_(dst, src)
            self.copies_in += len(src)

    # ---- overlapped path: launch a bucket as soon as all of its grads exist -------------------------
    def _make_hook(self, bi):
        def hook(param):
            left = self._ready.get(bi, len(self.buckets[bi][1])) - 1
            self._ready[bi] = left
            if left == 0:
                self._launch(bi)
        return hook

    def _reduce(self, flat):
        if self._avg_native:
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _launch(self, bi):
        flat = self.buckets[bi][0]
        self._adopt(bi)
        if self.transport == "p2p":
            cur = torch.cuda.current_stream(flat.device)
            if self._stream is not None:
                self._stream.wait_stream(cur)
                self._p2p[bi].allreduce_avg(self._stream)
            else:
                self._p2p[bi].allreduce_avg(cur)
            self._pending.append((bi, None))
            return
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                work = self._reduce(flat)
        else:
            work = self._reduce(flat)
        self._pending.append((bi, work))

    # ---- the call at the seam -------------------------------------------------------------------------
    def __call__(self, params=None):
        """Average the gradients over all ranks (in place).  No-op for a single process."""
        if self.world == 1:
            return
        launched = {bi for bi, _ in self._pending}
        for bi in range(len(self.buckets)):          # not overlapped, or buckets whose hooks did not all fire
            if bi not in launched:
                self._launch(bi)
        for bi, work in self._pending:
            if work is not None:
                work.wait()
            flat = self.buckets[bi][0]
            if self._stream is not None:
                torch.cuda.current_stream(flat.device).wait_stream(self._stream)
            if work is not None and not self._avg_native:
                flat.div_(self.world)
        self._pending.clear()
        self._ready.clear()

    finish = __call__

    def p2p_timeouts(self):
        """Barriers of the p2p transport that gave up waiting for a peer (must be 0; synchronises)."""
        return sum(pb.timeouts() for pb in self._p2p if pb is not None)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
        if any(pb is not None for pb in self._p2p):
            # the parameters' .grad are views of library-owned memory: detach them before it is freed
            for flat, ps, views in self.buckets:
                for p, v in zip(ps, views):
                    if p.grad is v:
                        p.grad = None
            self.buckets = []
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)      # nobody unmaps a bucket a peer's kernel may still touch
            for pb in self._p2p:
                if pb is not None:
                    pb.close()
            self._p2p = []


def attach_learner_allreduce(intrinsic_novelty, group=None, overlap=False, transport="nccl"):
    """Install reducers on the two intrinsic-novelty learners of this package (embedding + RND)."""
    emb = intrinsic_novelty.epi_novel.embedding
    rnd = intrinsic_novelty.ll_novel
    emb.grad_hook = GradAllReducer(list(emb.parameters()), group, overlap=overlap, transport=transport)
    rnd.grad_hook = GradAllReducer(list(rnd.predictor.parameters()), group, overlap=overlap, transport=transport)
    return emb.grad_hook, rnd.grad_hook


def attach_r2d2_allreduce(learner, group=None, overlap=True, transport="nccl"):
    """The third seam: ``R2D2Learner.step`` (ngu/models/r2d2/r2d2_learner.py:44-58).

    ``learner`` is the reference's ``R2D2Learner`` (or anything with its attributes: ``policy``, ``optimizer``,
    ``model_hypr``, ``update_count``, ``r2d2_loss_rms``, ``logger``).  Its ``step`` is replaced by the same
    statements with the gradient average inserted between ``loss.backward()`` (:53) and
    ``clip_grad_norm_`` (:54) - the patch a maintainer would make in place:

        -        self.optimizer.zero_grad()
        +        self.grad_reducer.zero_grad()
                 loss.backward()
        +        self.grad_reducer()                     # average over the data-parallel learners
                 nn.utils.clip_grad_norm_(self.policy.parameters(), self.model_hypr['adam_clip_norm'])

    With ``overlap=True`` the 32.75 MB bucket is split in two so that the head / LSTM gradients travel while the
    convolution trunk is still in backward.  Returns the reducer.
    """
    import torch.nn as nn
    params = list(learner.policy.parameters())
    learner.grad_reducer = GradAllReducer(params, group, bucket_bytes=(20 << 20) if overlap else (64 << 20), overlap=overlap,
                                          transport=transport)

    def step(self, td_errors, weights):
        loss = (td_errors.pow(2).mean(dim=0) * weights).mean()                       # r2d2_learner.py:51
        self.grad_reducer.zero_grad()                                                 # :52 (keeps the bucket views)
        loss.backward()                                                               # :53
        self.grad_reducer()                                                           # new: average over ranks
        nn.utils.clip_grad_norm_(self.policy.parameters(), self.model_hypr['adam_clip_norm'])   # :54
        self.optimizer.step()                                                         # :55
        self.update_count += 1                                                        # :56
        self.r2d2_loss_rms.update(loss.item())                                        # :57
        self.logger.log_scalar('R2D2Loss', self.r2d2_loss_rms.mean, self.update_count)   # :58

    learner.step = types.MethodType(step, learner)
    return learner.grad_reducer
