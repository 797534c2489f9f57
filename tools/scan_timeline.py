"""Scratch experiment: per-warp timeline of one scan launch (ramp, steady state, tail, merge), by SM."""
import ctypes as C, json, os, sys
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine

pct = lambda x: [round(float(v), 2) for v in np.percentile(x, [0, 10, 50, 90, 100])]
CASES = [(256, 30000, 0, d, "0.6,16") for d in (0, 1, 3, 7, 4, 32, 0)]
for B, N, variant, dbg, sched in CASES:
    os.environ["NGU_EPI_DEBUG_SCAN"] = str(dbg)
    os.environ["NGU_EPI_SCHED"] = sched
    os.environ["NGU_EPI_PDL"] = "0"
    eng = EpisodicEngine(B, N, device=0, scan_variant=variant)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    for _ in range(20): eng.step(q, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): eng.scan(q)
    e1.record(); torch.cuda.synchronize()
    scan_us = e0.elapsed_time(e1) * 10
    buf = np.zeros((148 * 4 * 32, 8), np.uint64)
    n = eng._lib.ngu_epi_debug_timeline(eng._h, buf.ctypes.data, buf.shape[0])
    t = buf[:n].astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :4] - t0) / 1e3
    smid = t[:, 4]
    out = dict(B=B, variant=variant, debug=dbg, sched=sched, scan_plus_ranksum_us=round(scan_us, 1), warps=int(len(t)),
               start_us=pct(rel[:, 0]), first_tile_us=pct(rel[:, 1]), scan_end_us=pct(rel[:, 2]),
               flush_end_us=pct(rel[:, 3]), flush_dur_us=pct(rel[:, 3] - rel[:, 2]), scan_dur_us=pct(rel[:, 2] - rel[:, 1]))
    # per-SM view: mean scan end of the warps on each SM; spread between and within SMs
    sm_end = np.array([rel[smid == s, 2].mean() for s in np.unique(smid)])
    sm_std_within = np.array([rel[smid == s, 2].std() for s in np.unique(smid)])
    out["sm_mean_end_us"] = pct(sm_end)
    out["within_sm_std_us"] = pct(sm_std_within)
    order = np.argsort(sm_end)
    tiles = t[:, 5]
    sm_tiles = np.array([tiles[smid == s].sum() for s in np.unique(smid)])
    out["last_run_tiles"] = pct(tiles)
    print(json.dumps(out), flush=True)
    if dbg == 0 and sched == "1.0,0":
        np.save("gpurun_out/timeline_b256.npy", np.concatenate([rel, smid[:, None]], axis=1))
    eng.close()
