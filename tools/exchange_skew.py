"""Where does the multi-GPU step lose time?  Per-rank %globaltimer stamps around the NVLink rank-sum exchange
(epilogue.cuh EPI_STAMP 2 -> 3 bracket p2p_allreduce) for ORDINARY (not pipelined) lock-step steps.

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/exchange_skew.py [--actors 256] [--capacity 30000]

Per rank and iteration (the last of `burst` back-to-back steps, then a synchronize to read the stamps):
  wait_us    = stamp3 - stamp2: this rank's own clock, no cross-GPU assumption.  The rank that arrives LAST waits only
               for the NVLink round (its own stores are the last to land everywhere); everybody else waits for it.
  arrive_us  = stamp2 relative to the earliest rank of that iteration, exit_us = stamp3 likewise.  The %globaltimer of
               different GPUs differ by constant offsets of up to a second (measured on the 8-GPU box); all ranks leave
               the exchange within one NVLink latency of each other in real time, so the mean difference of the exit
               stamps is taken as the offset and removed (the residual exit spread is printed as the sanity check).
Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

os.environ["NGU_EPI_DEBUG_EPILOGUE"] = "1"
os.environ["NGU_EPI_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def analyse(a):
    """a[rank, iteration, (stamp2, stamp3, ...)] in ns of each GPU's own %globaltimer (offsets between GPUs: up to a
    second, constant).  All ranks leave the exchange within one NVLink latency of each other in real time, so the mean
    difference of the exit stamps IS the clock offset; arrivals are compared after removing it."""
    a = (a - a[:, :, :2].min()).astype(np.float64)                   # int64 ns -> small exact float64
    s2, s3 = a[:, :, 0], a[:, :, 1]
    world = a.shape[0]
    wait = (s3 - s2) / 1e3
    offset = (s3 - s3[0:1]).mean(axis=1, keepdims=True)               # clock of rank r minus clock of rank 0
    arr = (s2 - offset) / 1e3
    arr = arr - arr.min(axis=0, keepdims=True)                        # us after the earliest rank of that iteration
    leave = (s3 - offset) / 1e3
    leave = leave - leave.min(axis=0, keepdims=True)
    last = arr.argmax(axis=0)
    r2 = lambda xs: [round(float(x), 2) for x in xs]
    return {
        "wait_us_mean_per_rank": r2(wait.mean(axis=1)), "wait_us_p95_per_rank": r2(np.percentile(wait, 95, axis=1)),
        "wait_us_min_over_ranks_mean": round(float(wait.min(axis=0).mean()), 2),
        "wait_us_max_over_ranks_mean": round(float(wait.max(axis=0).mean()), 2),
        "clock_offset_us_vs_rank0": r2(offset[:, 0] / 1e3),
        "arrive_us_after_earliest_mean_per_rank": r2(arr.mean(axis=1)),
        "arrival_spread_us_mean": round(float(arr.max(axis=0).mean()), 2),
        "arrival_spread_us_p95": round(float(np.percentile(arr.max(axis=0), 95)), 2),
        "exit_spread_us_mean": round(float(leave.max(axis=0).mean()), 2),
        "last_arriver_histogram": [int((last == r).sum()) for r in range(world)],
        "note": "wait = stamp3 - stamp2 on the rank's own clock (robust); arrivals compare clocks after removing the "
                "constant offsets estimated from the exit stamps (exit spread = residual of that calibration + "
                "NVLink latency differences)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--actors", type=int, default=256, help="actors PER GPU")
    ap.add_argument("--capacity", type=int, default=30000)
    ap.add_argument("--iters", type=int, default=150)
    ap.add_argument("--burst", type=int, default=4)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ngu_b200.sharding import ShardedEpisodicEngine

    B, N = args.actors, args.capacity
    sh = ShardedEpisodicEngine(B * world, N, device=local, exchange="p2p")
    eng = sh.engine
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    eng._memory.normal_(generator=g)
    qs = [torch.randn(B, 32, device="cuda", generator=g) for _ in range(8)]
    al = [1 + torch.randn(B, device="cuda", generator=g) for _ in range(8)]
    for i in range(600):                        # clocks ramped, peers mapped
        sh.step(qs[i % 8], al[i % 8])
    torch.cuda.synchronize()
    rows = np.zeros((args.iters, 4), np.int64)        # stamp2, stamp3, step start (stamp 8: epilogue entry), scan end
    st = np.zeros(10, np.uint64)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms = 0.0
    for it in range(args.iters):
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for j in range(args.burst):
            sh.step(qs[(it + j) % 8], al[(it + j) % 8])
        ev1.record()
        torch.cuda.synchronize()
        step_ms += ev0.elapsed_time(ev1) / args.burst
        assert eng._lib.ngu_epi_debug_timeline(eng._h, st.ctypes.data, -10) == 10
        s = st.astype(np.int64)
        rows[it] = (s[2], s[3], s[8], s[0])
    state = sh.get_state()
    assert state["exchange_timeouts"] == 0
    t = torch.from_numpy(rows).cuda()
    allrows = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allrows, t)
    ms = torch.tensor([step_ms / args.iters], device="cuda", dtype=torch.float64)   # (stamps travel as int64: 1.8e18 ns does not fit a double)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = analyse(torch.stack(allrows).cpu().numpy())         # [world, iters, 4] ns
        out.update({"n_gpus": world, "actors_per_gpu": B, "capacity": N, "iters": args.iters, "burst": args.burst,
                    "ordinary_step_us": round(float(ms.item()) * 1e3, 2)})
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
