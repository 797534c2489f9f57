"""Scratch: does keeping (part of) the memory stream resident in L2 across steps pay?  NGU_EPI_L2=mode,frac sweep
(ScanParams::l2_mode) over shard sizes; device-resident step time, interleaved repeats on one box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine

def run(B, N, l2, steps=400, warm=100):
    if l2 is None:
        os.environ.pop("NGU_EPI_L2", None)
    else:
        os.environ["NGU_EPI_L2"] = l2
    eng = EpisodicEngine(B, N, device=0)
    eng._memory.normal_()
    qs = [torch.randn(B, 32, device="cuda") for _ in range(4)]
    a = 1 + torch.randn(B, device="cuda")
    for i in range(warm): eng.step(qs[i % 4], a)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): eng.step(qs[i % 4], a)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    eng.scan_timing_begin(100)
    for i in range(100): eng.step(qs[i % 4], a)
    ms, n = eng.scan_timing_end()
    eng.close()
    return dict(B=B, N=N, MB=round(B * N * 128 / 1e6), l2=l2 or "0 (evict_first)", step_us=round(best, 2), scan_us=round(ms / n * 1e3, 2))

if __name__ == "__main__":
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(32, 30000), (64, 30000), (16, 30000), (64, 5000), (256, 30000)]
    for B, N in shapes:
        MB = B * N * 128 / 1e6
        fr = [f for f in (1.0, 0.75, 0.5, 0.375, 0.25, 0.125, 0.0625) if 20 <= f * MB <= 130] or [1.0]
        for l2 in [None, "1"] + [f"2,{f}" for f in fr] + [None]:
            print(json.dumps(run(B, N, l2)), flush=True)
