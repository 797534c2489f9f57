"""Scratch: scan-kernel chunk-schedule sweep (NGU_EPI_SCHED=static_frac,c2,c3,n3) on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine


def run(B, N, sched, variant=0, steps=200, warm=50):
    os.environ["NGU_EPI_SCHED"] = sched
    eng = EpisodicEngine(B, N, device=0, scan_variant=variant)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    for _ in range(warm): eng.step(q, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for _ in range(steps): eng.step(q, a)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    eng.scan_timing_begin(steps)
    for _ in range(steps): eng.step(q, a)
    ms, n = eng.scan_timing_end()
    g = eng.scan_geometry()
    eng.close()
    return dict(B=B, N=N, sched=sched, variant=variant, step_us=round(best, 2), scan_us=round(ms / n * 1e3, 2),
                scan_GBs=round(B * N * 128 / (ms / n) / 1e6, 0), s1=g["static_tiles_per_warp"], c2=g["dynamic_chunk_tiles"], chunks=g["chunks"])


if __name__ == "__main__":
    scheds = sys.argv[1:] or ["1.0,0", "0.8,8", "0.7,8", "0.6,8", "0.5,8", "0.4,8", "0.6,16", "0.6,12", "0.6,6", "0.7,8,4,2", "0.6,8,4,2", "0.6,16,4,4", "0.0,16"]
    for (B, N, v) in [(256, 30000, 0), (256, 30000, 4), (1024, 30000, 0)]:
        for sc in scheds:
            try:
                print(json.dumps(run(B, N, sc, v)), flush=True)
            except Exception as e:
                print("failed", B, N, sc, e, flush=True)
