"""Scratch experiment: where does the non-scan part of a step go? (not part of the product)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine


def run(B, N, pdl, phases, variant=0, steps=200, warm=30):
    os.environ["NGU_EPI_PDL"] = str(pdl)
    os.environ["NGU_EPI_DEBUG_PHASES"] = str(phases)
    eng = EpisodicEngine(B, N, device=0, scan_variant=variant)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda")
    a = 1 + torch.randn(B, device="cuda")
    for _ in range(warm): eng.step(q, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): eng.step(q, a)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / steps * 1e3
    eng.scan_timing_begin(steps)
    for _ in range(steps): eng.step(q, a)
    ms, n = eng.scan_timing_end()
    eng.close()
    return dict(B=B, N=N, variant=variant, pdl=pdl, phases=phases, step_us=round(us, 2), scan_us=round(ms / n * 1e3, 2),
                GBs=round(B * N * 128 / us / 1e3, 1))


if __name__ == "__main__":
    for (B, N) in [(64, 5000), (32, 30000), (1, 30000), (8, 30000), (256, 1000)]:
        for variant in (0,):
            for pdl, phases in [(1, 15), (1, 0)]:
                print(json.dumps(run(B, N, pdl, phases, variant)), flush=True)
