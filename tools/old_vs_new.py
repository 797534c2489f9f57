"""Scratch: device-resident step time, previous round's library (build/libold.so, ABI 2) vs the current one."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ngu_b200._cabi import NguEpiCfg
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def run(path, B, N, steps=300):
    lib = C.CDLL(path)
    lib.ngu_epi_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngu_epi_step.argtypes = [C.c_void_p] * 6
    lib.ngu_epi_destroy.argtypes = [C.c_void_p]
    cfg = NguEpiCfg(n_actors=B, capacity=N, dim=32, k=10, mode=0, device=0, kernel_epsilon=1e-4, cluster_distance=0.008,
                    pseudo_counts=1e-3, max_similarity=8.0, max_reward_scale=5.0, reserved0=0.0, prior_count=10.0,
                    scan_ctas_per_sm=0, scan_variant=0)
    mem = torch.randn(B, N, 32, device="cuda")
    h = C.c_void_p()
    assert lib.ngu_epi_create(C.byref(cfg), mem.data_ptr(), C.byref(h)) == 0
    q = torch.randn(B, 32, device="cuda"); a = 1 + torch.randn(B, device="cuda")
    r = torch.empty(B, device="cuda"); r2 = torch.empty(B, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def step(): assert lib.ngu_epi_step(h, q.data_ptr(), a.data_ptr(), r.data_ptr(), r2.data_ptr(), st) == 0
    for _ in range(50): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(steps): step()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    lib.ngu_epi_destroy(h)
    return round(best, 2)

for B, N in [(1, 30000), (8, 30000), (32, 30000), (64, 5000), (256, 1000), (256, 5000), (256, 30000)]:
    print(json.dumps(dict(B=B, N=N, old_us=run(os.path.join(ROOT, "build", "libold.so"), B, N),
                          new_us=run(os.path.join(ROOT, "ngu_b200", "lib", "libngu_epi.so"), B, N))), flush=True)
