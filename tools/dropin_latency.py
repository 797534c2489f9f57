"""End-to-end latency of the drop-in IntrinsicNovelty.compute_intrinsic_novelty with the real CNN
producers: eager vs CUDA-graph replay (scratch measurement)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngu_b200.pytorch_util as ptu
from ngu_b200.engine import DEFAULT_HYPR
from ngu_b200.intrinsic_novelty import IntrinsicNovelty


class L:
    def log_scalar(self, *a): pass


def run(B, N, graph, steps=100, pinned=False):
    ptu.init_device()
    inn = IntrinsicNovelty(B, 18, (1, 84, 84), dict(DEFAULT_HYPR, episodic_memory_capacity=N), L()).to(ptu.device)
    inn.epi_novel._ensure_engine()._memory.normal_()
    if graph:
        inn.enable_cuda_graph()
    obs = [torch.randn(B, 1, 84, 84).clamp_(-5, 5) for _ in range(4)]       # pageable, as an env wrapper hands them over
    if pinned:
        obs = [o.pin_memory() for o in obs]
    for i in range(10):
        inn.compute_intrinsic_novelty(obs[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        inn.compute_intrinsic_novelty(obs[i % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dict(B=B, N=N, graph=graph, pinned_obs=pinned, ms_per_step=round(dt * 1e3, 4), qps=round(B / dt))


if __name__ == "__main__":
    for B, N in [(256, 30000), (64, 5000), (1, 30000)]:
        for graph, pinned in ((False, False), (False, True), (True, False)):
            print(json.dumps(run(B, N, graph, pinned=pinned)), flush=True)
