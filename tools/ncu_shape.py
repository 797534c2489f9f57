"""Tiny driver for ncu captures of one shape: python tools/ncu_shape.py B N [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine
B, N = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
eng = EpisodicEngine(B, N, device=0); eng._memory.normal_()
qs = [torch.randn(B, 32, device="cuda") for _ in range(4)]
a = 1 + torch.randn(B, device="cuda")
for i in range(steps): eng.step(qs[i % 4], a)
torch.cuda.synchronize()
print("done", eng.scan_geometry())
