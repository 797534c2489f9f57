"""Scratch: what does the merge job of a single-actor scan spend its time on?"""
import json, os, sys
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
os.environ["NGU_EPI_PDL"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ngu_b200.engine import EpisodicEngine
for B, N in [(1, 30000), (64, 5000)]:
    os.environ['NGU_EPI_SCHED'] = '0.6,16,24,8'
    eng = EpisodicEngine(B, N, device=0)
    eng._memory.normal_()
    q = torch.randn(B, 32, device="cuda")
    for _ in range(5): eng.scan(q)
    torch.cuda.synchronize()
    buf = np.zeros((148 * 4 * 32, 8), np.uint64)
    # zero the device timeline first so that stale job records do not confuse us: run once more after reading
    n = eng._lib.ngu_epi_debug_timeline(eng._h, buf.ctypes.data, buf.shape[0])
    t = buf[:n].astype(np.int64)
    t0 = t[t[:, 0] > 0][:, 0].min()
    tt = t[t[:, 0] > 0]
    print('warps', len(tt), 'start', [round(float(v - t0) / 1e3, 2) for v in np.percentile(tt[:, 0], [0, 50, 100])], 'first_tile', [round(float(v - t0) / 1e3, 2) for v in np.percentile(tt[:, 1], [0, 50, 100])],
          'scan_end', [round(float(v - t0) / 1e3, 2) for v in np.percentile(tt[:, 2], [0, 50, 100])], 'warp_end', [round(float(v - t0) / 1e3, 2) for v in np.percentile(tt[:, 3], [0, 50, 100])])
    jobs = t[(t[:, 5] >= t0) & (t[:, 7] >= t[:, 5])]
    jobs = jobs[np.argsort(jobs[:, 7])]
    print('jobs', len(jobs), 'job_end pct', [round(float(v - t0) / 1e3, 2) for v in np.percentile(jobs[:, 7], [0, 50, 90, 100])], 'scan_end max', round(float(t[t[:,0]>0][:, 2].max() - t0) / 1e3, 2))
    for r in list(jobs[:3]) + list(jobs[-8:]):
        print(json.dumps(dict(B=B, N=N, scan_end=round((r[2]-t0)/1e3, 2), job_start=round((r[5]-t0)/1e3, 2), job_end=round((r[7]-t0)/1e3, 2),
                              warp_end=round((r[3]-t0)/1e3, 2), cand=int(r[6] & 0xFFFFFFFF),                              )), flush=True)
    eng.close()
