"""Scratch: where does a small-shape step go?  Per-warp scan stamps + epilogue phase stamps of one step,
relative to the scan's first warp start, plus event-timed launch-to-launch step time."""
import json, os, sys
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
os.environ["NGU_EPI_DEBUG_EPILOGUE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine

pct = lambda x: [round(float(v), 2) for v in np.percentile(x, [0, 50, 100])]
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(1, 30000), (64, 5000), (8, 30000), (256, 1000)]
for B, N in SHAPES:
    eng = EpisodicEngine(B, N, device=0); eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    for _ in range(50): eng.step(q, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): eng.step(q, a)
    e1.record(); torch.cuda.synchronize()
    step_us = e0.elapsed_time(e1) * 5
    acc = []
    for it in range(20):
        for _ in range(3): eng.step(q, a)
        torch.cuda.synchronize()
        st = np.zeros(10, np.uint64)
        assert eng._lib.ngu_epi_debug_timeline(eng._h, st.ctypes.data, -10) == 10
        tl = np.zeros((148 * 4 * 32, 8), np.uint64)
        nw = eng._lib.ngu_epi_debug_timeline(eng._h, tl.ctypes.data, tl.shape[0])
        t = tl[:nw].astype(np.int64); t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        s = st.astype(np.int64)
        acc.append([ (t[:, 0].max() - t0), (t[:, 1].min() - t0), (t[:, 1].max() - t0), (t[:, 2].min() - t0), (t[:, 2].max() - t0),
                     (t[:, 3].max() - t0), (s[8] - t0), (s[0] - t0), (s[7] - t0)])
    c = t[t[:, 5] > 0]                                  # static schedules: the CTA combiners of the last step
    comb = dict(combiners=int(len(c)), start_after_own_stream_end=pct((c[:, 5] - c[:, 2]) / 1e3), level1_us=pct((c[:, 6] - c[:, 5]) / 1e3),
                level2_us=pct((c[:, 7] - c[:, 6]) / 1e3), exit_after_level2=pct((c[:, 3] - c[:, 7]) / 1e3),
                last_combiner_start=round(float((c[:, 5].max() - t0) / 1e3), 2)) if len(c) else {}
    m = np.mean(np.array(acc), axis=0) / 1e3
    names = ["last_warp_start", "first_tile_min", "first_tile_max", "stream_end_min", "stream_end_max", "flush_merge_end_max",
             "epilogue_entry", "epilogue_wait_return", "epilogue_last_stamp"]
    print(json.dumps(dict(B=B, N=N, warps=int(len(t)), step_us=round(step_us, 2), geometry=eng.scan_geometry(),
                          **{n: round(float(v), 2) for n, v in zip(names, m)}, **comb)), flush=True)
    eng.close()
