"""A/B of two builds of libngu_epi.so on one box, interleaved child processes: ordinary / pipelined device steps and the
synchronous host call.  Usage: lib_ab.py build/libprev.so ngu_b200/lib/libngu_epi.so [BxN ...]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(path, shapes):
    import ngu_b200.build as b
    b.LIB_PATH = path
    import ngu_b200._cabi as c
    c.LIB_PATH = path
    import numpy as np, torch
    from ngu_b200.engine import EpisodicEngine
    out = {}
    for B, N in shapes:
        eng = EpisodicEngine(B, N, device=0)
        eng._memory.normal_()
        qs = [torch.randn(B, 32, device="cuda") for _ in range(8)]
        a = 1 + torch.randn(B, device="cuda")
        qh = [q.cpu().pin_memory() for q in qs]; ah = a.cpu().pin_memory()
        steps = 2000 if B * N < 4e6 else 500
        for i in range(1200 if B * N > 4e6 else 300): eng.step(qs[i % 8], a, pipelined=True)
        torch.cuda.synchronize()
        r = {}
        for mode in ("ordinary", "pipelined"):
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps): eng.step(qs[i % 8], a, pipelined=(mode == "pipelined"))
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / steps * 1e3)
            r[mode] = round(best, 2)
        ro = np.empty((2, B), np.float32)
        for i in range(30): eng.step_host(qh[i % 8], ah, out=ro)
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for i in range(steps // 2): eng.step_host(qh[i % 8], ah, out=ro)
            best = min(best, (time.perf_counter() - t0) / (steps // 2) * 1e6)
        r["host_call"] = round(best, 2)
        assert eng.get_state()["exchange_timeouts"] == 0
        out[f"{B}x{N}"] = r
        eng.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], [tuple(int(v) for v in s.split("x")) for s in sys.argv[3:]])
    else:
        libs = [os.path.abspath(p) for p in sys.argv[1:3]]
        shapes = sys.argv[3:] or ["256x30000", "32x30000", "64x5000"]
        for rep in range(2):
            for p in libs:
                res = subprocess.run([sys.executable, __file__, "--child", p] + shapes, capture_output=True, text=True)
                line = [l for l in res.stdout.splitlines() if l.startswith("{")]
                print(json.dumps({"lib": os.path.relpath(p, ROOT), "rep": rep, **(json.loads(line[-1]) if line else {"error": res.stderr[-400:]})}), flush=True)
