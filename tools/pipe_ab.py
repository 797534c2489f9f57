"""Device-resident step time, ordinary (ngu_epi_step) vs software-pipelined (ngu_epi_step_pipelined) launches,
interleaved repeats on one box.  Usage: pipe_ab.py [BxN ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine

def run(eng, qs, a, pipelined, steps):
    for i in range(50): eng.step(qs[i % len(qs)], a, pipelined=pipelined)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): eng.step(qs[i % len(qs)], a, pipelined=pipelined)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    return best

if __name__ == "__main__":
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]] or \
        [(1, 30000), (64, 5000), (8, 30000), (16, 30000), (32, 30000), (64, 30000), (128, 30000), (256, 30000)]
    for B, N in shapes:
        eng = EpisodicEngine(B, N, device=0)
        eng._memory.normal_()
        qs = [torch.randn(B, 32, device="cuda") for _ in range(8)]
        a = 1 + torch.randn(B, device="cuda")
        steps = 2000 if B * N < 4e6 else 600
        for i in range(300): eng.step(qs[i % 8], a)
        res = {}
        for rep in range(2):
            for mode in (False, True):
                res.setdefault(mode, []).append(round(run(eng, qs, a, mode, steps), 2))
        st = eng.get_state()
        bytes_ = (128 * N + 340) * B
        print(json.dumps(dict(B=B, N=N, ordinary_us=res[False], pipelined_us=res[True],
                              hbm_floor_us=round(bytes_ / 6556.5e9 * 1e6, 2),
                              pipelined_TBs=round(bytes_ / min(res[True]) / 1e6, 3), timeouts=st["exchange_timeouts"])), flush=True)
        eng.close()
