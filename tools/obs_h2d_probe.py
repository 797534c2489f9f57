"""Scratch: how should the drop-in move a PAGEABLE (B,1,84,84) float observation batch to the GPU?
(the caller hands CPU tensors to compute_intrinsic_novelty every step.)"""
import json, sys, time
import torch

def bench(fn, n=60):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

for B in (256, 64, 8):
    obs = [torch.randn(B, 1, 84, 84) for _ in range(4)]            # pageable, as an env wrapper produces them
    obs_pin = [o.pin_memory() for o in obs]
    dev = torch.empty(B, 1, 84, 84, device="cuda")
    pin = torch.empty(B, 1, 84, 84).pin_memory()
    i = [0]
    def pageable(): i[0] += 1; dev.copy_(obs[i[0] % 4], non_blocking=True)
    def pinned_src(): i[0] += 1; dev.copy_(obs_pin[i[0] % 4], non_blocking=True)
    def staged(): i[0] += 1; pin.copy_(obs[i[0] % 4]); dev.copy_(pin, non_blocking=True)
    out = dict(B=B, MB=round(B * 84 * 84 * 4 / 1e6, 2), pageable_ms=round(bench(pageable), 4), pinned_source_ms=round(bench(pinned_src), 4),
               staged_ms=round(bench(staged), 4))
    for chunks in (2, 4, 8):
        if B < chunks: continue
        bounds = [B * c // chunks for c in range(chunks + 1)]
        def chunked():
            i[0] += 1; o = obs[i[0] % 4]
            for c in range(chunks):
                lo, hi = bounds[c], bounds[c + 1]
                pin[lo:hi].copy_(o[lo:hi]); dev[lo:hi].copy_(pin[lo:hi], non_blocking=True)
        out[f"chunked{chunks}_ms"] = round(bench(chunked), 4)
    torch.set_num_threads(8)
    print(json.dumps(out), flush=True)
