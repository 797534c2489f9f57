"""Scratch: time the scan variants on the GPU box (not part of the product)."""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine

def time_variant(B, N, variant, ctas=0, steps=200, warm=30):
    eng = EpisodicEngine(B, N, device=0, scan_variant=variant, scan_ctas_per_sm=ctas)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda")
    for _ in range(warm): eng.scan(q)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps): eng.scan(q)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    # full step
    for _ in range(warm): eng.step(q)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps): eng.step(q)
    ev[1].record(); torch.cuda.synchronize()
    ms2 = ev[0].elapsed_time(ev[1]) / steps
    gb = B * N * 128 / 1e9
    g = eng.scan_geometry()
    eng.close()
    return dict(B=B, N=N, variant=variant, ctas=ctas, scan_ms=round(ms, 4), scan_GBs=round(gb / ms * 1e3, 1), step_ms=round(ms2, 4), step_GBs=round(gb / ms2 * 1e3, 1), **g)

if __name__ == "__main__":
    for (B, N) in [(256, 30000), (1024, 30000)]:
        for v in range(8):
            try:
                print(json.dumps(time_variant(B, N, v)), flush=True)
            except Exception as e:
                print("variant", v, "failed:", e, flush=True)
