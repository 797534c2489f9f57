"""Scratch: cost of reset_memory_if_done (intrinsic_novelty.py:36-40) on the device, 256 x 30k."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine
B, N = 256, 30000
eng = EpisodicEngine(B, N, device=0); eng._memory.normal_()
for n_done in (0, 1, 4, 26, 256):
    done = torch.zeros(B, dtype=torch.bool, device="cuda"); done[:n_done] = True
    for _ in range(5): eng.reset(done)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): eng.reset(done)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(json.dumps(dict(n_done=n_done, us=round(us, 2), GBs=round(n_done * N * 128 / us / 1e3, 1))), flush=True)
eng.close()
