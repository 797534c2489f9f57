import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine
B, N = 256, 30000
dyn = EpisodicEngine(B, N, device=0)
os.environ["NGU_EPI_SCHED"] = "1.0,0"
sta = EpisodicEngine(B, N, device=0)
g = torch.Generator(device="cuda").manual_seed(5)
dyn._memory.normal_(generator=g); sta._memory.copy_(dyn._memory)
t0 = time.time(); bad = 0
for t in range(40000):
    q = torch.randn(B, 32, device="cuda", generator=g); a = 1 + 2 * torch.randn(B, device="cuda", generator=g)
    r1, _ = dyn.step(q, a); r2, _ = sta.step(q, a)
    if t % 250 == 0:
        if not torch.equal(r1, r2): bad += 1; print("MISMATCH at", t, flush=True)
    if t % 1013 == 7:
        done = torch.rand(B, device="cuda", generator=g) < 0.03
        dyn.reset(done); sta.reset(done)
print("steps 40000 mismatches", bad, "memory equal", bool(torch.equal(dyn._memory, sta._memory)), "timeouts", dyn.get_state()["exchange_timeouts"], sta.get_state()["exchange_timeouts"], "secs", round(time.time() - t0, 1))
