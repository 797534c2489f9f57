"""Learner gradient all-reduce hook (BASELINE north_star "the learner's
embedding/RND/R2D2 gradient allreduce runs over NCCL on NVLink").

The reference has NO gradient averaging anywhere: this is new, data-parallel
functionality attached at the three seams where gradients exist and clipping has not happened yet:

    ngu/models/intrinsic_novelty/episodic_novelty/embedding.py:60-61      (182,866 params, 0.73 MB)
    ngu/models/intrinsic_novelty/lifelong_novelty/lifelong_novelty.py:64-65 (473,376 params, 1.89 MB)
    ngu/models/r2d2/r2d2_learner.py:53-54                                  (8,188,595 params, 32.75 MB)

``GradAllReducer`` flattens the gradients into a few contiguous buckets and averages them with
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).  With
``overlap=True`` the buckets are reduced on a side stream as soon as autograd has produced all of
their gradients (post-accumulate-grad hooks, last layers first), so the transfer of the 32 MB R2D2
bucket hides behind the rest of backward; ``finish()`` — the call placed at the seam — only waits.

Usage at a seam (what ``Embedding.step`` / ``LifelongNovelty.step`` in this package do through their
``grad_hook`` attribute, and what a maintainer adds to ``R2D2Learner.step``):

    reducer = GradAllReducer(params)           # once
    loss.backward()
    reducer(params)                             # <- new line: average over ranks
    nn.utils.clip_grad_norm_(params, 40)
    optimizer.step()
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params, group=None, bucket_bytes=32 << 20, overlap=False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
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
        if self.overlap:
            dev = self.params[0].device
            self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
            for bi, (_, ps, _) in enumerate(self.buckets):
                for p in ps:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _close(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
        views, off = [], 0
        for p in ps:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.buckets.append((flat, ps, views))

    @property
    def nbytes(self):
        return sum(f.numel() * f.element_size() for f, _, _ in self.buckets)

    # ---- overlapped path: launch a bucket as soon as all of its grads exist -------------------------
    def _make_hook(self, bi):
        def hook(param):
            left = self._ready.get(bi, len(self.buckets[bi][1])) - 1
            self._ready[bi] = left
            if left == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        flat, ps, views = self.buckets[bi]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                torch._foreach_copy_(views, [p.grad for p in ps])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            torch._foreach_copy_(views, [p.grad for p in ps])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((bi, work))

    # ---- the call at the seam -------------------------------------------------------------------------
    def __call__(self, params=None):
        """Average the gradients over all ranks (in place).  No-op for a single process."""
        if self.world == 1:
            return
        if self.overlap:
            launched = {bi for bi, _ in self._pending}
            for bi in range(len(self.buckets)):      # buckets whose hooks did not all fire (unused params)
                if bi not in launched and all(p.grad is not None for p in self.buckets[bi][1]):
                    self._launch(bi)
        else:
            for bi, (_, ps, _) in enumerate(self.buckets):
                if all(p.grad is not None for p in ps):
                    self._launch(bi)
        for bi, work in self._pending:
            work.wait()
            flat, ps, views = self.buckets[bi]
            if self._stream is not None:
                torch.cuda.current_stream(flat.device).wait_stream(self._stream)
            flat.div_(self.world)
            torch._foreach_copy_([p.grad for p in ps], views)
        self._pending.clear()
        self._ready.clear()

    finish = __call__

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def attach_learner_allreduce(intrinsic_novelty, group=None, overlap=False):
    """Install reducers on the two intrinsic-novelty learners of this package (embedding + RND)."""
    emb = intrinsic_novelty.epi_novel.embedding
    rnd = intrinsic_novelty.ll_novel
    emb.grad_hook = GradAllReducer(list(emb.parameters()), group, overlap=overlap)
    rnd.grad_hook = GradAllReducer(list(rnd.predictor.parameters()), group, overlap=overlap)
    return emb.grad_hook, rnd.grad_hook
