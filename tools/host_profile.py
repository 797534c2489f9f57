import os, sys, time
os.environ["NGU_EPI_HOST_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(256, 30000), (64, 5000)]
for B, N in SHAPES:
    eng = EpisodicEngine(B, N, device=0); eng._memory.normal_()
    rng = np.random.default_rng(0)
    q = rng.standard_normal((B, 32), dtype=np.float32); a = (1 + rng.standard_normal(B)).astype(np.float32)
    for _ in range(100): eng.step_host(q, a)
    t0 = time.perf_counter()
    for _ in range(500): eng.step_host(q, a)
    print(B, N, "python-level us/step", round((time.perf_counter() - t0) / 500 * 1e6, 2), flush=True)
    # device-resident for comparison
    qd = torch.from_numpy(q).cuda(); ad = torch.from_numpy(a).cuda()
    for _ in range(50): eng.step(qd, ad)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): eng.step(qd, ad)
    e1.record(); torch.cuda.synchronize()
    print(B, N, "device us/step", round(e0.elapsed_time(e1) / 300 * 1e3, 2), flush=True)
    eng.close()
