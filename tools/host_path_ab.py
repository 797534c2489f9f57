"""Scratch: A/B of the host-buffer step (ngu_epi_step_host): inputs staged by the copy engine vs pulled by a PDL-chained kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine

KNOB = sys.argv[1] if len(sys.argv) > 1 else "NGU_EPI_HOST_STAGE"
MODES = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("copy", "kernel")


def run(B, N, mode, steps=300):
    os.environ[KNOB] = mode
    eng = EpisodicEngine(B, N, device=0)
    eng._memory.normal_()
    rng = np.random.default_rng(0)
    qs = [rng.standard_normal((B, 32), dtype=np.float32) for _ in range(8)]
    al = [(1 + rng.standard_normal(B)).astype(np.float32) for _ in range(8)]
    for i in range(50): eng.step_host(qs[i % 8], al[i % 8])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(steps): eng.step_host(qs[i % 8], al[i % 8])
        best = min(best, (time.perf_counter() - t0) / steps * 1e6)
    r = eng.step_host(qs[0], al[0])[0]
    eng.close()
    return dict(B=B, N=N, mode=mode, us_per_step=round(best, 2), qps=round(B / best * 1e6), r0=float(r[0]))

if __name__ == "__main__":
    for (B, N) in [(256, 30000), (64, 5000), (1, 30000), (32, 30000), (1024, 30000)]:
        for mode in MODES * 2:
            print(json.dumps(run(B, N, mode)), flush=True)
