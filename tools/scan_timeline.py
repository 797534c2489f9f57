"""Scratch experiment: per-warp timeline of one scan launch (ramp, steady state, tail, merge)."""
import ctypes as C, json, os, sys
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine

for B, N, variant in [(256, 30000, 0), (256, 30000, 4), (1024, 30000, 0)]:
    eng = EpisodicEngine(B, N, device=0, scan_variant=variant)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    for _ in range(20): eng.step(q, a)
    torch.cuda.synchronize()
    buf = np.zeros((148 * 4 * 32, 4), np.uint64)
    n = eng._lib.ngu_epi_debug_timeline(eng._h, buf.ctypes.data, buf.shape[0])
    t = buf[:n].astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    pct = lambda x: [round(float(v), 2) for v in np.percentile(x, [0, 10, 50, 90, 100])]
    print(json.dumps(dict(B=B, variant=variant, warps=int(len(t)),
                          start_us=pct(rel[:, 0]), first_tile_us=pct(rel[:, 1]), scan_end_us=pct(rel[:, 2]),
                          flush_end_us=pct(rel[:, 3]), flush_dur_us=pct(rel[:, 3] - rel[:, 2]),
                          scan_dur_us=pct(rel[:, 2] - rel[:, 1]))), flush=True)
    # per-SM-ish view: end time by CTA
    eng.close()
