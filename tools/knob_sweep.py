"""Scratch: device-resident step time for a list of environment-knob settings, interleaved repeats on one box.
  python tools/knob_sweep.py B N REPEATS KEY=v1;KEY2=v2 KEY=v3 ...      ("-" = no knob)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200.engine import EpisodicEngine

def run(B, N, env, steps=600):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = EpisodicEngine(B, N, device=0)
    finally:
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
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
    for i in range(50): eng.step(qs[i % 8], al[i % 8], pipelined=True)
    e0.record()
    for i in range(steps): eng.step(qs[i % 8], al[i % 8], pipelined=True)
    e1.record(); torch.cuda.synchronize()
    pipe_us = e0.elapsed_time(e1) / steps * 1e3
    eng.scan_timing_begin(200)
    for i in range(200): eng.step(qs[i % 8], al[i % 8])
    ms, n = eng.scan_timing_end()
    assert eng.get_state()["exchange_timeouts"] == 0
    # host-buffer path (ngu_epi_step_host), synchronous per step
    import time, numpy as np
    rng = np.random.default_rng(0)
    qh = [rng.standard_normal((B, 32), dtype=np.float32) for _ in range(8)]
    ah = [(1 + rng.standard_normal(B)).astype(np.float32) for _ in range(8)]
    for i in range(50): eng.step_host(qh[i % 8], ah[i % 8])
    host = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(300): eng.step_host(qh[i % 8], ah[i % 8])
        host = min(host, (time.perf_counter() - t0) / 300 * 1e6)
    eng.close()
    return dict(B=B, N=N, env=env, step_us=round(us, 2), pipelined_us=round(pipe_us, 2), scan_alone_us=round(ms / n * 1e3, 2), host_step_us=round(host, 2))

if __name__ == "__main__":
    B, N, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    cases = [dict(kv.split("=", 1) for kv in a.split(";")) if a != "-" else {} for a in sys.argv[4:]]
    for rep in range(reps):
        for env in cases:
            print(json.dumps(run(B, N, env)), flush=True)
