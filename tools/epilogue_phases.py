"""Scratch: where does the epilogue kernel's post-wait time go? (%globaltimer stamps at phase boundaries)"""
import json, os, sys
os.environ["NGU_EPI_DEBUG_EPILOGUE"] = "1"
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine
names = ["wait->loaded", "loaded->ranksums", "ranksums->(exchange)", "finish:rms", "finish:rewards", "log sums", "append"]
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(256, 30000), (64, 5000), (1, 30000)]
for B, N in SHAPES:
    eng = EpisodicEngine(B, N, device=0); eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    acc = np.zeros(7); n = 0; tail = 0.0; entry = 0.0; prewait = 0.0
    for it in range(30):
        for _ in range(5): eng.step(q, a)
        torch.cuda.synchronize()
        st = np.zeros(10, np.uint64)
        assert eng._lib.ngu_epi_debug_timeline(eng._h, st.ctypes.data, -10) == 10
        tl = np.zeros((148 * 4 * 32, 8), np.uint64)
        nw = eng._lib.ngu_epi_debug_timeline(eng._h, tl.ctypes.data, tl.shape[0])
        scan_done = tl[:nw, 3].max()
        s = st.astype(np.int64)
        acc += np.diff(s[:8]) / 1e3; tail += (s[0] - int(scan_done)) / 1e3; n += 1
        entry += (int(scan_done) - s[8]) / 1e3; prewait += (s[9] - s[8]) / 1e3
    print(json.dumps(dict(B=B, N=N, scan_last_warp_to_epilogue_wait_return_us=round(tail / n, 2),
                          epilogue_resident_before_scan_end_us=round(entry / n, 2), prewait_work_us=round(prewait / n, 2),
                          **{k: round(float(v), 2) for k, v in zip(names, acc / n)}, total_post_wait_us=round(float(acc.sum() / n), 2))), flush=True)
    eng.close()
