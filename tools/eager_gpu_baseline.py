"""Secondary baseline: the reference's episodic-novelty op sequence as it runs when
ptu.device is CUDA - generic ATen kernels through PyTorch eager (episodic_novelty.py:51-78: broadcast sub,
pow, sum, sqrt, topk, host round trip for the running mean, ~10 tiny elementwise ops, row write) - timed on
the same B200 as bench.py.  Self-contained restatement (no oracle import); numpy mirrors mpi_moments."""
import json, sys, time
import numpy as np, torch

def run(B, N, k=10, steps=30, warm=5):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    mem = torch.randn(N, B, 32, device=dev, generator=g)          # the reference's slot-major layout
    qs = [torch.randn(B, 32, device=dev, generator=g) for _ in range(4)]
    mean, count, idx = np.zeros(k), float(k), 0
    def step(q):
        nonlocal mean, count, idx
        euc = ((mem - q.unsqueeze(0)) ** 2).sum(dim=-1).sqrt()      # :51-52
        kn = euc.topk(k, dim=0)                                     # :54
        x = kn.values.t().cpu().numpy()                             # :56 (device sync)
        bm = x.sum(axis=0, dtype=np.float32) / np.float32(B)
        mean = mean + (bm - mean) * B / (count + B); count += B
        nd = kn.values / torch.as_tensor(mean, dtype=torch.float32, device=dev).unsqueeze(-1)   # :58
        nd = torch.max(nd - 0.008, torch.zeros_like(nd))            # :60-65
        kv = 1e-4 / (nd + 1e-4)                                     # :67-68
        s = kv.sum(dim=0).sqrt() + 0.001                            # :71-72
        r = 1.0 / s
        r[s > 8] = 0.0                                              # :75-78
        mem[idx] = q; idx = (idx + 1) % N                           # :80,86-89
        return r.cpu()                                              # :84
    for i in range(warm): step(qs[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps): step(qs[i % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dict(B=B, N=N, ms_per_step=round(dt * 1e3, 3), queries_per_s=round(B / dt), peak_mem_GB=round(torch.cuda.max_memory_allocated() / 1e9, 2))

if __name__ == "__main__":
    for B, N in [(256, 30000), (64, 5000), (1, 30000)]:
        print(json.dumps(dict(torch_eager_gpu=run(B, N))), flush=True)
