"""Scratch: step time at 256 x 30k for tapered dynamic schedules (NGU_EPI_TAPER=half_share,quarter_share), interleaved repeats."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine

def run(B, N, taper, sched=None, steps=600):
    os.environ["NGU_EPI_TAPER"] = taper
    if sched: os.environ["NGU_EPI_SCHED"] = sched
    else: os.environ.pop("NGU_EPI_SCHED", None)
    eng = EpisodicEngine(B, N, device=0)
    eng._memory.normal_()
    g = torch.Generator(device="cuda").manual_seed(1)
    qs = [torch.randn(B, 32, device="cuda", generator=g) for _ in range(8)]
    al = [1 + torch.randn(B, device="cuda", generator=g) for _ in range(8)]
    for i in range(400): eng.step(qs[i % 8], al[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): eng.step(qs[i % 8], al[i % 8])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / steps * 1e3
    assert eng.get_state()["exchange_timeouts"] == 0
    geo = eng.scan_geometry()
    eng.close()
    return dict(B=B, N=N, taper=taper, sched=sched, step_us=round(us, 2), chunks=geo["chunks"])

if __name__ == "__main__":
    B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 30000
    cases = [("0,0", None), ("0.2,0.1", None), ("0.3,0.2", None), ("0.0,0.2", None), ("0.4,0.0", None), ("0.2,0.3", None),
             ("0.3,0.2", "0.7,16"), ("0.2,0.1", "0.5,32"), ("0.3,0.3", "0.6,32")]
    for rep in range(2):
        for taper, sched in cases:
            print(json.dumps(run(B, N, taper, sched)), flush=True)
