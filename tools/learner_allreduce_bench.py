"""Learner gradient all-reduce micro-benchmark (BASELINE config 5): the three gradient
sets of the reference's learners (0.73 MB embedding, 1.89 MB RND predictor, 32.75 MB R2D2) averaged
over the ranks with ngu_b200.learner_allreduce.GradAllReducer (NCCL over NVLink).

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/learner_allreduce_bench.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ngu_b200.learner_allreduce import GradAllReducer

SETS = {"embedding": 182866, "rnd_predictor": 473376, "r2d2": 8188595}   # parameter counts, SURVEY.md 2.2


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    for name, n in SETS.items():
        # a plausible layer structure: chunks of <= 1.6M elements (the LSTM weight of the reference is 3188x2048)
        sizes, left = [], n
        while left > 0:
            s = min(left, 1_600_000)
            sizes.append(s); left -= s
        params = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in sizes]
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1))
        red = GradAllReducer(params)
        for _ in range(5):
            red(params)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 50
        e0.record()
        for _ in range(iters):
            red(params)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = n * 4
        out[name] = {"bytes": nbytes, "ms": round(ms.item(), 4),
                     "algbw_GBs": round(nbytes / ms.item() / 1e6, 1),
                     "busbw_GBs": round(nbytes / ms.item() / 1e6 * 2 * (world - 1) / world, 1)}
        # correctness: mean of (rank+1) over ranks, applied repeatedly stays at that fixed point after the 1st call
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1))
        red(params)
        want = (world + 1) / 2
        assert all(torch.allclose(p.grad, torch.full_like(p, want)) for p in params)
    if rank == 0:
        print(json.dumps({"learner_allreduce": out, "world": world}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
