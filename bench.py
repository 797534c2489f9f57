#!/usr/bin/env python
"""bench.py — episodic intrinsic-reward queries/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path, config 3 (weak)
    python bench.py --config {1,2,3,5} [--scaling weak|strong]        # the other SURVEY.md 8d configurations
    python bench.py --impl reference [--gpus N --steps K --warmup W] # the reference's own CPU path
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # one rank per GPU

A step = one pass of the hot path over one batch of synthetic controllable states: distance scan of every
actor's whole memory + top-k(10) + running-mean update + reward epilogue (x RND multiplier clipped to
[1,5]) + append.

Default workload (BASELINE.json configs[2]): 256 actors x 30 000-slot memory x D=32
per GPU, memory fully populated with N(0,1) rows (983 MB per GPU, 7.8x the 126 MB L2, so every step
streams from HBM).  Multi-GPU: actors are sharded, 256 per GPU (weak scaling, the headline); the only
exchange is the rank-sum exchange fused into the epilogue kernel over NVLink peer memory.

One JSON line on stdout (rank 0).
  value      whole-job queries/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-buffer C-ABI call (ngu_epi_step_host: H2D of the step's states +
             modulators, D2H of the rewards, synchronous)
  roofline   scan kernel: algorithmic bytes / live CUDA-event launch time vs the measured HBM peak
  parity     in-run correctness evidence: the very data plane that is timed (N>1: NVLink p2p exchange),
             checked against the unsharded CPU oracle BEFORE the timed region, plus post-run health
             (no exchange timeouts, replicated running mean bit-identical on all ranks).  A failure
             exits non-zero.
  strong     (N>1) config 3 exactly as SURVEY.md 8d writes it: 256 actors IN TOTAL, 256/N per GPU
  secondary  (N=1) config 2 (64 x 5k, L2-resident, flagged), the config-1 substitute (B=1 drop-in with the
             real CNNs vs the imported reference), and the torch-eager GPU baseline (the reference module
             itself with ptu.device = cuda)
  cpu_baseline  the imported, unmodified reference (oracle/_ref) on this box's host cores
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its
# version banner on fd 1, the reference prints "No GPU..."), so the real stdout is kept aside and fd 1
# is pointed at stderr.
RESULT_OUT = None


def claim_stdout():
    global RESULT_OUT
    if RESULT_OUT is None:
        sys.stdout.flush()
        RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


ACTORS_PER_GPU = 256
CAPACITY = 30000
K_NEIGHBORS = 10
METRIC = "episodic intrinsic-reward queries/sec"
UNIT = "queries/s"
CLOCK_RAMP_STEPS = 1500          # untimed steps after the W warm-up steps (fixed count: same collectives on every rank)


def bytes_per_query(capacity):
    """memory read + query + append + outputs."""
    return 128 * capacity + 340


def workload_name(actors, capacity):
    cap = f"{capacity // 1000}k" if capacity % 1000 == 0 else str(capacity)
    return f"synthetic episodic kNN: {actors} actors/GPU x {cap}-slot memory x D=32, k={K_NEIGHBORS}"


def l2_note(B, N):
    mb = B * N * 128 / 1e6
    if mb > 2 * 126:
        return f"working set {mb:.0f} MB per GPU vs 126 MB L2 (inputs larger than L2)"
    if mb > 100:
        return f"working set {mb:.0f} MB per GPU ~ the 126 MB L2: partially cache-resident, flagged"
    return f"working set {mb:.0f} MB per GPU fits the 126 MB L2: cache-resident, flagged"


def base_config(B, N, world, exchange):
    """The config keys BOTH arms emit (the driver compares them)."""
    return {"workload": workload_name(B, N), "actors_per_gpu": B, "global_actors": B * world, "capacity": N,
            "k": K_NEIGHBORS, "mode": "reference", "parallelism": f"actor-shard x{world}", "l2": l2_note(B, N),
            "exchange": exchange}


def exchange_note(world, mode):
    if world == 1:
        return "none"
    if mode == "off":
        return "OFF (timing experiment: ranks independent)"
    if mode == "p2p":
        return "21-double rank sums + 5 log-only sums per step, NVLink peer stores fused into the epilogue kernel"
    return "21-double rank-sum all-reduce (torch.distributed) per step"


# ---------------------------------------------------------------------------------
# clocks (NVML, sampled from a thread while the timed region runs)
# ---------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None

    def _one(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._one()
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self._one()
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._one()
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------
# CPU arm: the imported, unmodified reference (oracle/_ref); the torch port only if it is not staged
# ---------------------------------------------------------------------------------
def _versions():
    import torch
    return {"torch": torch.__version__, "numpy": np.__version__}


class CpuArm:
    """The reference's CPU implementation of the path on `actors` x `capacity` (identity embedding, injected
    alpha: SURVEY.md section 8c).  kind 'reference' = the imported reference module, 'port' = oracle/torch_port.py."""

    def __init__(self, actors, capacity, threads):
        import torch
        from oracle import ref_runner
        self.torch = torch
        self.B, self.N = actors, capacity
        torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(1234)
        if ref_runner.available():
            self.kind = "reference"
            self.impl = ("imported unmodified reference: ngu.models.intrinsic_novelty.IntrinsicNovelty."
                         "compute_intrinsic_novelty (oracle/_ref, identity embedding, injected alpha)")
            self.r = ref_runner.ReferenceRunner(actors, capacity, device="cpu")
            self.r.fill_memory(1234)
            self._step = lambda q, a: self.r.step(q, a)
        else:
            from oracle.torch_port import TorchCpuPort
            self.kind = "port"
            self.impl = "torch-CPU op-for-op port of the reference (oracle/torch_port.py); oracle/_ref not staged"
            port = TorchCpuPort(actors, capacity, k=K_NEIGHBORS)
            port.memory = torch.randn(capacity, actors, 32, generator=g)
            self._step = lambda q, a: port.step(q, a)
        self.qs = [torch.randn(actors, 32, generator=g) for _ in range(4)]
        self.al = [(1.0 + torch.randn(actors, generator=g)).double().numpy() for _ in range(4)]

    def run(self, steps, warmup, threads=None):
        """(queries/s, ms/step) over exactly `steps` timed steps after `warmup` untimed ones."""
        if threads:
            self.torch.set_num_threads(threads)
        for i in range(warmup):
            self._step(self.qs[i % 4], self.al[i % 4])
        t0 = time.perf_counter()
        for i in range(steps):
            self._step(self.qs[i % 4], self.al[i % 4])
        dt = time.perf_counter() - t0
        return self.B * steps / dt, dt / steps * 1e3


def cpu_sample_actors(steps, warmup, full_actors, budget_s=150.0, est_ms_per_actor_step=0.5):
    """Bounded sample: the reference's time is linear in the actor count (a streaming scan per actor), so when
    K+W full-size steps would not end within a few minutes the sample keeps the memory depth and shrinks the
    actor batch."""
    est = (steps + warmup) * full_actors * est_ms_per_actor_step / 1e3
    if est <= budget_s:
        return full_actors
    return max(8, int(full_actors * budget_s / est) // 8 * 8)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from oracle.ref_runner import cpu_model_string
    B_full, N = args.actors_per_gpu, args.capacity
    threads = os.cpu_count() or 1
    B = cpu_sample_actors(args.steps, args.warmup, B_full)
    arm = CpuArm(B, N, threads)
    qps, ms = arm.run(args.steps, args.warmup)
    sample = (f"{args.steps} timed + {args.warmup} warm-up steps of {B} actors x {N} slots"
              + ("" if B == B_full else f" (bounded sample of the {B_full}-actor batch: same memory depth, fewer actors)")
              + f", {arm.impl}, {threads} threads")
    cfg = base_config(B_full, N, args.gpus, exchange_note(args.gpus, "p2p"))     # the very keys the GPU arm emits
    detail = {"device": "host CPU", "cpu_model": cpu_model_string(), "versions": _versions(), "threads": threads,
              "note": "one CPU job on rank 0 whatever --gpus says: a rate for the ratio, not a scaled run"}
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "detail": detail,
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": arm.kind, "sample": sample,
                         "cpu_model": cpu_model_string()},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_leg(B, N, budget_s=14.0):
    """cpu_baseline of the GPU line (N=1, rank 0): all threads and one thread, bounded."""
    from oracle.ref_runner import cpu_model_string
    threads = os.cpu_count() or 1
    arm = CpuArm(B, N, threads)
    t0 = time.perf_counter()
    arm.run(1, 1)                                        # calibrate: one warm-up + one step
    per = max((time.perf_counter() - t0) / 2, 1e-3)
    steps = int(max(5, min(40, budget_s * 0.6 / per)))
    qps, ms = arm.run(steps, 0)
    one = CpuArm(max(8, B // 16), N, 1)                  # 1 thread: a 1/16 actor sample of the same depth
    q1, ms1 = one.run(3, 1, threads=1)
    return {"value": qps, "unit": UNIT, "cores": threads, "kind": arm.kind,
            "sample": f"{steps} steps (median-free mean after 2 warm-ups) of the same {B} x {N} workload, {arm.impl}, "
                      f"{ms:.1f} ms/step",
            "one_thread": {"value": q1, "unit": UNIT, "sample": f"3 steps of {one.B} actors x {N} slots, 1 thread, {ms1:.1f} ms/step"},
            "cpu_model": cpu_model_string(), "versions": _versions()}


# ---------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the scan kernel from the committed `ncu --set full` capture (not this run)."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get("dram_bytes_per_launch"), f"profiles/scan_traffic.json ({d.get('source', 'ncu --set full capture')})"
        except Exception:
            pass
    return None, None


class Dist:
    """torch.distributed plumbing of one rank."""

    def __init__(self, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_equal_bytes(self, arr: np.ndarray) -> bool:
        """True iff `arr` is bit-identical on every rank."""
        if self.world == 1:
            return True
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).astype(np.int32)).to(self.dev)
        lo, hi = t.clone(), t.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
        return bool(self.torch.equal(lo, hi))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_sharded(B_local, N, world, local_rank, args):
    from ngu_b200.sharding import ShardedEpisodicEngine
    return ShardedEpisodicEngine(B_local * world, N, device=local_rank, k=K_NEIGHBORS, scan_variant=args.scan_variant,
                                 track_log_stats=True, exchange=os.environ.get("NGU_BENCH_EXCHANGE", "auto"))


def measure_shape(D, sh, B, N, rank, local_rank, K, W, ramp, want_clocks=True, e2e_steps=None):
    """Timed region + roofline pass + e2e for one (B per GPU, N) shape on an existing sharded engine."""
    torch = D.torch
    eng, dev, world = sh.engine, D.dev, D.world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    eng._memory.normal_(generator=g)                    # fully populated memory
    NQ = 8
    qs = [torch.randn(B, 32, device=dev, generator=g) for _ in range(NQ)]
    al = [1.0 + torch.randn(B, device=dev, generator=g) for _ in range(NQ)]
    q_host = [q.cpu().pin_memory() for q in qs]
    a_host = [a.cpu().pin_memory() for a in al]

    # Device-resident steps are enqueued back to back from buffers that are never rewritten: exactly the contract of
    # the software-pipelined launch (ngu_epi_step_pipelined, include/ngu_epi.h) - same results bit for bit, the next
    # scan streams under the previous step's merge tail / epilogue / NVLink exchange.  NGU_BENCH_PIPELINE=0: ordinary.
    pipe = os.environ.get("NGU_BENCH_PIPELINE", "1") != "0" and (world == 1 or sh.exchange == "p2p")
    for i in range(W + ramp):
        sh.step(qs[i % NQ], al[i % NQ], pipelined=pipe)
        if i % 128 == 127:
            torch.cuda.synchronize()
    D.barrier()

    # ---- timed region: K steps, device-resident inputs, CUDA events on the launching stream ----
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank) if want_clocks else None
    if clocks:
        clocks.__enter__()
    D.barrier()
    ev0.record()
    for i in range(K):
        sh.step(qs[i % NQ], al[i % NQ], pipelined=pipe)
    ev1.record()
    D.barrier()
    if clocks:
        clocks.__exit__()
    ms_total = D.max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - launches0
    # the same K steps launched the ordinary way (every scan waits for the previous epilogue): reported next to `value`
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    ev2.record()
    for i in range(K):
        sh.step(qs[i % NQ], al[i % NQ])
    ev3.record()
    D.barrier()
    ms_unpiped = D.max_over_ranks(ev2.elapsed_time(ev3))
    # roofline pass: the same K steps again with a CUDA-event pair around every launch of the scan kernel
    # (kept out of the region above: the event records break the programmatic-dependent-launch overlap)
    # (at most 40 launches: on the power-capped boxes of this pool a long run of isolated, un-overlapped launches drifts
    # with the clocks - 155-159 us per launch over 20-50 launches, 174 us over 200 at 1.6 GHz - while the timed region
    # above, which is what `value` reports, does not)
    KR = min(K, 40)
    eng.scan_timing_begin(KR)
    for i in range(KR):
        sh.step(qs[i % NQ], al[i % NQ])
    scan_ms, scan_calls = eng.scan_timing_end()

    # ---- e2e: host buffers through the C ABI (H2D + step + D2H, synchronous) ----
    host_call = world == 1 or sh.exchange == "p2p"   # the fused single-call step (exchange inside the epilogue kernel)

    r_out = [np.empty((2, B), np.float32) for _ in range(2)]             # the caller's result buffers (rewards land here)

    def e2e_step(i):
        if host_call:
            return eng.step_host(q_host[i % NQ], a_host[i % NQ], out=r_out[i & 1])[0]   # pinned CPU tensors, passed by address
        q = q_host[i % NQ].to(dev, non_blocking=True)
        a = a_host[i % NQ].to(dev, non_blocking=True)
        return sh.step(q, a)[0].cpu()
    KE = e2e_steps or K
    for i in range(max(3, min(W, 20))):
        e2e_step(i)
    D.barrier()
    t0 = time.perf_counter()
    for i in range(KE):
        e2e_step(i)
    e2e_local = time.perf_counter() - t0
    D.barrier()
    e2e_s = D.max_over_ranks(e2e_local)

    scan_avg_ms = scan_ms / max(scan_calls, 1)
    achieved = bytes_per_query(N) * B / (scan_avg_ms / 1e3) / 1e9
    return {
        "value": B * world * K / (ms_total / 1e3), "ms_per_step": ms_total / K, "pipelined": pipe,
        "ms_per_step_unpipelined": ms_unpiped / K,
        "e2e_value": B * world * KE / e2e_s, "e2e_ms_per_call": e2e_s / KE * 1e3, "e2e_steps": KE,
        "host_call": host_call, "launches": int(launches), "clocks": clocks.summary() if clocks else None,
        "scan_launch_ms": scan_avg_ms, "achieved_gbs": achieved, "qs": qs, "NQ": NQ, "gen": g,
    }


def stream_only_ms(B, N, local_rank, args, qs, gen):
    """Diagnostic: the SAME scan kernel with its arithmetic, top-k and flushes switched off (results invalid), on a
    second memory of the same size: what this TMA access pattern alone takes on this box.  Never a throughput."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    os.environ["NGU_EPI_DEBUG_SCAN"] = "15"
    try:
        bare = EpisodicEngine(B, N, device=local_rank, k=K_NEIGHBORS, scan_variant=args.scan_variant)
    finally:
        del os.environ["NGU_EPI_DEBUG_SCAN"]
    bare._memory.normal_(generator=gen)
    for i in range(30):
        bare.scan(qs[i % len(qs)])
    torch.cuda.synchronize()
    bare.scan_timing_begin(100)
    for i in range(100):
        bare.scan(qs[i % len(qs)])
    ms, n = bare.scan_timing_end()
    # the same bare stream launched as software-pipelined steps (what `value` is timed with): the practical read-only
    # ceiling of this access pattern when the stream never drains between launches
    al = 1.0 + torch.zeros(B, device=qs[0].device)
    for i in range(30):
        bare.step(qs[i % len(qs)], al, pipelined=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        bare.step(qs[i % len(qs)], al, pipelined=True)
    e1.record()
    torch.cuda.synchronize()
    piped = e0.elapsed_time(e1) / 100
    bare.close()
    return ms / max(n, 1), piped


# ---------------------------------------------------------------------------------
# in-run parity: the data plane that is timed, against the unsharded CPU oracle (checker only)
# ---------------------------------------------------------------------------------
def parity_phase(D, rank, local_rank, args):
    """Before the timed region.  N>1: the NVLink p2p exchange (what `value` and `e2e` run on) with ragged actor
    shards, resets and the alpha clip; then a 300-step collect_sequence-style loop of BASELINE config 4 (256
    actors sharded over the ranks, done with p = 1/1000, L = 5, ring wrap).  N=1: the same two cases unsharded.
    Bars: kNN slots and d^2 bit-exact, rewards <= 1e-5 relative (BASELINE.json north_star)."""
    torch, world = D.torch, D.world
    from ngu_b200.sharding import ShardedEpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    exchange = os.environ.get("NGU_BENCH_EXCHANGE", "auto")
    out = {"cases": [], "bars": "kNN slots + d^2 bit-exact, rewards <= 1e-5 relative, vs oracle/episodic_oracle.py (CPU, unsharded)"}
    worst = 0.0
    fails = []
    for name, B, N, steps, p_done, host_every in (("ragged_shards_37x3000", 37, 3000, 10, 0.2, 2),
                                                  ("config4_loop_256x200_300steps", 256, 200, 300, 1e-3, 3)):
        sh = ShardedEpisodicEngine(B, N, device=local_rank, k=K_NEIGHBORS, exchange=exchange)
        rng = np.random.default_rng(2024)            # same stream on every rank
        mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
        mem0[rng.random(N) < 0.2] = 0.0
        sh.engine.load_memory(mem0[:, sh.lo:sh.hi])
        orc = EpisodicOracle(B, N, k=K_NEIGHBORS, use_c_scan=True)
        orc.memory[...] = mem0
        case_worst = 0.0
        host_ok = world == 1 or sh.exchange == "p2p"
        for t in range(steps):
            q = rng.standard_normal((B, 32), dtype=np.float32)
            alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)      # clips at 1 and at L = 5
            done = rng.random(B) < p_done
            want = orc.step(q, alpha)
            if host_ok and t % host_every == 0:       # the host-buffer call (what e2e times) ...
                r = sh.engine.step_host(q[sh.lo:sh.hi], alpha[sh.lo:sh.hi])[0]
            else:                                     # ... and the device-pointer call (what value times)
                r = sh.step(torch.from_numpy(q[sh.lo:sh.hi]).to(D.dev), torch.from_numpy(alpha[sh.lo:sh.hi]).to(D.dev))[0].cpu().numpy()
            d2, idx = sh.engine.topk()
            if not np.array_equal(idx.T, want["idx"][:, sh.lo:sh.hi]):
                fails.append(f"{name} step {t} rank {rank}: kNN slots differ")
            if not np.array_equal(d2.T.view(np.uint32), want["d2"][:, sh.lo:sh.hi].view(np.uint32)):
                fails.append(f"{name} step {t} rank {rank}: d^2 bits differ")
            w = want["reward"][sh.lo:sh.hi]
            rel = float((np.abs(r - w) / np.maximum(np.abs(w), 1e-30)).max())
            case_worst = max(case_worst, rel)
            if not rel <= 1e-5:
                fails.append(f"{name} step {t} rank {rank}: reward rel err {rel:.2e}")
            orc.reset(done)
            if done[sh.lo:sh.hi].any() or t % 2:
                sh.reset(torch.from_numpy(done[sh.lo:sh.hi]))
            if len(fails) > 5:
                break
        st = sh.get_state()                           # collective in p2p mode (flushes the one-step-late log-only sums)
        if not np.allclose(st["ed_mean"], orc.ed_rms.mean, rtol=1e-6, atol=0):
            fails.append(f"{name} rank {rank}: running mean differs from the oracle")
        if st["ed_count"] != float(orc.ed_rms.count):
            fails.append(f"{name} rank {rank}: ed_count {st['ed_count']} != {float(orc.ed_rms.count)}")
        if abs(st["epi_count"] - (1e-4 + steps * B)) > 1e-6 * steps * B:
            fails.append(f"{name} rank {rank}: log-only tracker saw {st['epi_count']} samples, expected {steps * B}")
        if not np.array_equal(sh.engine.dump_memory(), orc.memory[:, sh.lo:sh.hi]):
            fails.append(f"{name} rank {rank}: episodic memory differs from the oracle")
        if not D.all_equal_bytes(st["ed_mean"]):
            fails.append(f"{name}: replicated running mean not bit-identical across ranks")
        if st["exchange_timeouts"] != 0:
            fails.append(f"{name} rank {rank}: exchange_timeouts = {st['exchange_timeouts']}")
        case_worst = D.max_over_ranks(case_worst)
        worst = max(worst, case_worst)
        out["cases"].append({"name": name, "actors": B, "capacity": N, "steps": steps, "exchange": sh.exchange,
                             "worst_rel": case_worst})
        sh.engine.close()
    # ---- what `value` times: bursts of software-pipelined steps, enqueued without any host synchronisation ----
    name, B, N = "pipelined_bursts_96x2500", 96, 2500
    sh = ShardedEpisodicEngine(B, N, device=local_rank, k=K_NEIGHBORS, exchange=exchange)
    if world == 1 or sh.exchange == "p2p":
        rng = np.random.default_rng(4048)
        mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
        sh.engine.load_memory(mem0[:, sh.lo:sh.hi])
        orc = EpisodicOracle(B, N, k=K_NEIGHBORS, use_c_scan=True)
        orc.memory[...] = mem0
        case_worst, total = 0.0, 0
        for steps in (2, 5, 40):
            qh = rng.standard_normal((steps, B, 32), dtype=np.float32)
            ah = (1.0 + 2.0 * rng.standard_normal((steps, B))).astype(np.float32)
            qd = torch.from_numpy(qh[:, sh.lo:sh.hi].copy()).to(D.dev)
            ad = torch.from_numpy(ah[:, sh.lo:sh.hi].copy()).to(D.dev)
            r = torch.empty((steps, sh.hi - sh.lo), dtype=torch.float32, device=D.dev)
            e = torch.empty_like(r)
            torch.cuda.synchronize()
            for t in range(steps):
                sh.step(qd[t], ad[t], pipelined=True, out=(r[t], e[t]))
            torch.cuda.synchronize()
            got = r.cpu().numpy()
            for t in range(steps):
                want = orc.step(qh[t], ah[t])
                w = want["reward"][sh.lo:sh.hi]
                rel = float((np.abs(got[t] - w) / np.maximum(np.abs(w), 1e-30)).max())
                case_worst = max(case_worst, rel)
                if not rel <= 1e-5:
                    fails.append(f"{name} burst of {steps}, step {t}, rank {rank}: reward rel err {rel:.2e}")
            d2, idx = sh.engine.topk()                # the burst's last step, where the pipeline is deepest
            if not np.array_equal(idx.T, want["idx"][:, sh.lo:sh.hi]) or \
                    not np.array_equal(d2.T.view(np.uint32), want["d2"][:, sh.lo:sh.hi].view(np.uint32)):
                fails.append(f"{name} burst of {steps}, rank {rank}: kNN of the last step differs")
            total += steps
        st = sh.get_state()
        if not np.array_equal(sh.engine.dump_memory(), orc.memory[:, sh.lo:sh.hi]):
            fails.append(f"{name} rank {rank}: episodic memory differs from the oracle")
        if st["exchange_timeouts"] != 0 or st["cursor"] != orc.cursor:
            fails.append(f"{name} rank {rank}: timeouts {st['exchange_timeouts']}, cursor {st['cursor']} vs {orc.cursor}")
        case_worst = D.max_over_ranks(case_worst)
        worst = max(worst, case_worst)
        out["cases"].append({"name": name, "actors": B, "capacity": N, "steps": total, "exchange": sh.exchange,
                             "launch": "ngu_epi_step_pipelined, bursts of 2 / 5 / 40 steps without host sync", "worst_rel": case_worst})
    sh.engine.close()
    n_fail = D.max_over_ranks(float(len(fails)))
    key = "p2p_vs_oracle" if world > 1 else "single_gpu_vs_oracle"
    out[key] = "ok" if n_fail == 0 else "FAILED"
    out["worst_rel"] = worst
    if fails:
        out["failures"] = fails[:6]
    return out, n_fail == 0


def post_run_health(D, sh, parity):
    """After the timed region: no merge / exchange round timed out, and the replicated running mean is
    bit-identical on every rank (a silent timeout would sum zeros in: a FASTER wrong step)."""
    st = sh.get_state()
    ok = True
    parity["exchange_timeouts"] = int(D.max_over_ranks(float(st["exchange_timeouts"])))
    parity["ed_mean_identical_across_ranks"] = bool(D.all_equal_bytes(st["ed_mean"]))
    parity["ed_count_after_run"] = st["ed_count"]
    if parity["exchange_timeouts"] != 0 or not parity["ed_mean_identical_across_ranks"]:
        ok = False
    if not np.isfinite(st["ed_mean"]).all():
        ok = False
        parity["ed_mean_finite"] = False
    return ok


# ---------------------------------------------------------------------------------
# secondary measurements (N = 1)
# ---------------------------------------------------------------------------------
def gpu_eager_baseline(B, N, steps=20, warm=3):
    """The reference module ITSELF with ptu.device = cuda (torch-eager ATen kernels; identity embedding, injected
    alpha): what minoring/NGU runs when it is given this GPU - 'the kernel to beat on the same box'."""
    import torch
    from oracle import ref_runner
    if not ref_runner.available():
        return {"unavailable": "oracle/_ref not staged"}
    r = ref_runner.ReferenceRunner(B, N, device="cuda")
    r.fill_memory(1234)
    g = torch.Generator().manual_seed(7)
    qs = [torch.randn(B, 32, generator=g) for _ in range(4)]
    al = [(1.0 + torch.randn(B, generator=g)).double().numpy() for _ in range(4)]
    torch.cuda.reset_peak_memory_stats()
    for i in range(warm):
        r.step(qs[i % 4], al[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        r.step(qs[i % 4], al[i % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    out = {"value": B / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": steps,
           "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2),
           "what": "imported reference IntrinsicNovelty.compute_intrinsic_novelty with ptu.device = cuda (torch eager), "
                   "host tensors in / out like the e2e call"}
    del r
    torch.cuda.empty_cache()
    # put the reference's module-level device back on the CPU for whatever runs next
    ref_runner.import_reference(use_gpu=False)
    return out


def config1_substitute(local_rank, steps=200):
    """SURVEY.md 8d config 1 substitute: the full compute_intrinsic_novelty call with the real Embedding + RND CNNs at
    B = 1, N = 30 000, synthetic observations ~ N(0,1) clipped to +-5 - this repo's drop-in (eager and CUDA-graph
    replay) next to the imported reference on the host CPU."""
    import torch
    import ngu_b200.pytorch_util as ptu
    from ngu_b200.engine import DEFAULT_HYPR
    from ngu_b200.intrinsic_novelty import IntrinsicNovelty
    from oracle import ref_runner

    class _L:
        def log_scalar(self, *a, **k):
            pass
    B, N = 1, 30000
    g = torch.Generator().manual_seed(1)
    obs = [torch.randn(B, 1, 84, 84, generator=g).clamp_(-5, 5) for _ in range(8)]
    out = {"workload": "config-1 substitute: full compute_intrinsic_novelty, real CNNs, B=1, N=30000, obs ~ N(0,1) clipped to +-5"}
    ptu.init_device(index=local_rank)
    for graph in (False, True):
        inn = IntrinsicNovelty(B, 18, (1, 84, 84), dict(DEFAULT_HYPR, episodic_memory_capacity=N), _L()).to(ptu.device)
        inn.epi_novel._ensure_engine()._memory.normal_()
        if graph:
            inn.enable_cuda_graph()
        for i in range(10):
            inn.compute_intrinsic_novelty(obs[i % 8])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            inn.compute_intrinsic_novelty(obs[i % 8])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        out["cuda_graph" if graph else "eager"] = {"value": B / dt, "unit": UNIT, "ms_per_call": dt * 1e3, "steps": steps}
        inn.epi_novel._engine.close()
        del inn
    if ref_runner.available():
        r = ref_runner.ReferenceRunner(B, N, device="cpu", real_cnn=True, threads=os.cpu_count() or 1)
        r.fill_memory(1234)
        for i in range(2):
            r.step(obs[i % 8])
        n = 20
        t0 = time.perf_counter()
        for i in range(n):
            r.step(obs[i % 8])
        dt = (time.perf_counter() - t0) / n
        out["reference_cpu"] = {"value": B / dt, "unit": UNIT, "ms_per_call": dt * 1e3, "steps": n,
                                "what": "imported reference, real CNNs, host CPU, all threads"}
    return out


def learner_allreduce_leg(D, rank):
    """north_star: the learner's embedding / RND / R2D2 gradient all-reduce over NCCL on NVLink:
    the three gradient sets of the reference's learners averaged with ngu_b200.learner_allreduce.GradAllReducer."""
    torch, dist, world = D.torch, D.dist, D.world
    from ngu_b200.learner_allreduce import GradAllReducer
    sets = {"embedding": 182866, "rnd_predictor": 473376, "r2d2": 8188595}     # parameter counts, SURVEY.md 2.2
    out = {}
    cases = [(name, n, "nccl") for name, n in sets.items()]
    if world <= 8:
        cases += [(name + "_p2p", n, "p2p") for name, n in sets.items()]       # the library's own NVLink kernel (include/ngu_ar.h)
    for name, n, transport in cases:
        sizes, left = [], n
        while left > 0:
            s = min(left, 1_600_000)
            sizes.append(s)
            left -= s
        params = [torch.nn.Parameter(torch.zeros(s, device=D.dev)) for s in sizes]
        red = GradAllReducer(params, transport=transport)
        for p in params:
            p.grad.fill_(float(rank + 1))
        for _ in range(5):
            red(params)
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 50
        e0.record()
        for _ in range(iters):
            red(params)
        e1.record()
        torch.cuda.synchronize()
        ms = D.max_over_ranks(e0.elapsed_time(e1) / iters)
        for p in params:
            p.grad.fill_(float(rank + 1))
        red(params)
        want = (world + 1) / 2
        ok = all(torch.allclose(p.grad, torch.full_like(p, want)) for p in params)
        nbytes = n * 4
        out[name] = {"bytes": nbytes, "ms": round(ms, 4), "algbw_GBs": round(nbytes / ms / 1e6, 1),
                     "busbw_GBs": round(nbytes / ms / 1e6 * 2 * (world - 1) / world, 1), "average_correct": bool(ok),
                     "transport": transport}
        red.close()
    return out


# ---------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------
def run_gpu_arm(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        # no fallback on purpose: this arm measures the CUDA library and nothing else
        raise SystemExit("bench.py: no CUDA device - the product path is libngu_epi.so on a B200 and has no CPU fallback "
                         "(the CPU numbers are `--impl reference`)")
    torch.cuda.set_device(local_rank)
    D = Dist(world, local_rank)
    K, W = args.steps, args.warmup
    ok = True

    parity, p_ok = ({"skipped": "--no-parity"}, True) if args.no_parity else parity_phase(D, rank, local_rank, args)
    ok &= p_ok
    if rank == 0:
        log(f"parity: {json.dumps(parity)}")

    # ---- primary measurement ----
    N = args.capacity
    if args.scaling == "strong":
        if args.global_actors % world:
            raise SystemExit(f"--scaling strong: {args.global_actors} actors do not split evenly over {world} GPUs")
        B = args.global_actors // world
    else:
        B = args.actors_per_gpu
    sh = make_sharded(B, N, world, local_rank, args)
    m = measure_shape(D, sh, B, N, rank, local_rank, K, W, CLOCK_RAMP_STEPS)
    ok &= post_run_health(D, sh, parity)
    geo = sh.engine.scan_geometry()
    exchange_mode = sh.exchange

    stream_ms = stream_piped_ms = None
    if rank == 0 and not args.no_stream_only:
        try:
            stream_ms, stream_piped_ms = stream_only_ms(B, N, local_rank, args, m["qs"], m["gen"])
        except Exception as e:                      # a diagnostic must never cost the bench line
            log(f"stream-only diagnostic skipped: {e}")
    sh.engine.close()
    del sh
    torch.cuda.empty_cache()

    # ---- N > 1: config 3 as SURVEY.md 8d writes it (256 actors in total), and the learner all-reduce ----
    strong = None
    learner = None
    if world > 1 and args.scaling == "weak" and not args.no_secondary and args.global_actors % world == 0:
        Bs = args.global_actors // world
        shs = make_sharded(Bs, N, world, local_rank, args)
        ms_ = measure_shape(D, shs, Bs, N, rank, local_rank, max(K, 200), W, 300, want_clocks=False)
        ok &= post_run_health(D, shs, {})
        peak, _ = measured_peak()
        strong = {"scaling": "strong", "global_actors": args.global_actors, "actors_per_gpu": Bs,
                  "value": ms_["value"], "ms_per_step": ms_["ms_per_step"], "steps": max(K, 200),
                  "ms_per_step_unpipelined": ms_["ms_per_step_unpipelined"],
                  "e2e": {"value": ms_["e2e_value"], "ms_per_call": ms_["e2e_ms_per_call"]},
                  "roofline_ms_per_step": bytes_per_query(N) * Bs / (peak * 1e9) * 1e3,
                  "frac_of_roofline": bytes_per_query(N) * Bs / (ms_["ms_per_step"] / 1e3) / 1e9 / peak,
                  "scan_launch_ms": ms_["scan_launch_ms"], "l2": l2_note(Bs, N),
                  "note": "whole step (scan + merge + exchange + epilogue) against the HBM time of the shard; "
                          "speed-up over N=1 is this value / the N=1 line's value"}
        shs.engine.close()
        del shs
        torch.cuda.empty_cache()
    if world > 1 and not args.no_secondary:
        try:
            learner = learner_allreduce_leg(D, rank)
        except Exception as e:
            learner = {"error": str(e)[:200]}

    # ---- N = 1: the other SURVEY.md 8d configurations, same run ----
    secondary = None
    if world == 1 and rank == 0 and not args.no_secondary:
        secondary = {}
        try:
            B2, N2 = 64, 5000
            sh2 = make_sharded(B2, N2, 1, local_rank, args)
            m2 = measure_shape(D, sh2, B2, N2, rank, local_rank, 2000, 20, 300, want_clocks=False)
            peak, _ = measured_peak()
            secondary["config2_64x5k"] = {
                "workload": workload_name(B2, N2), "l2": l2_note(B2, N2), "value": m2["value"], "ms_per_step": m2["ms_per_step"],
                "ms_per_step_unpipelined": m2["ms_per_step_unpipelined"],
                "e2e": {"value": m2["e2e_value"], "ms_per_call": m2["e2e_ms_per_call"]}, "steps": 2000,
                "roofline": {"achieved": m2["achieved_gbs"], "peak": peak, "frac": m2["achieved_gbs"] / peak,
                             "launch_ms": m2["scan_launch_ms"], "note": "L2-resident set: the HBM peak is not the bound here (latency is)"}}
            sh2.engine.close()
            del sh2
        except Exception as e:
            secondary["config2_64x5k"] = {"error": str(e)[:200]}
        try:
            secondary["config1_b1_dropin"] = config1_substitute(local_rank)
        except Exception as e:
            secondary["config1_b1_dropin"] = {"error": str(e)[:200]}
        try:
            secondary["gpu_eager_baseline"] = gpu_eager_baseline(B, N)
        except Exception as e:
            secondary["gpu_eager_baseline"] = {"error": str(e)[:200]}
        torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = ncu_traffic()
        if (B, N) != (ACTORS_PER_GPU, CAPACITY):
            traffic, traffic_src = None, None          # the committed capture is of the 256 x 30k launch
        cfg = base_config(B, N, world, exchange_note(world, exchange_mode))
        detail = {"scan_geometry": geo, "clock_ramp_steps": CLOCK_RAMP_STEPS,
                  "launch_mode": ("software-pipelined steps (ngu_epi_step_pipelined): the scan of step t+1 streams under the merge "
                                  "tail / epilogue / exchange of step t; results identical to ngu_epi_step (parity object)")
                                 if m["pipelined"] else "ordinary steps (ngu_epi_step)",
                  "ms_per_step_unpipelined": m["ms_per_step_unpipelined"],
                  "warmup_note": f"{W} warm-up steps as asked + {CLOCK_RAMP_STEPS} further untimed steps so the SM clocks "
                                 "have ramped (fixed count: every rank issues the same sequence)"}
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "warmup_total": W + CLOCK_RAMP_STEPS, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "detail": detail,
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "ms_per_call": m["e2e_ms_per_call"],
                    "h2d_bytes_per_step": B * 32 * 4 + B * 4, "d2h_bytes_per_step": 2 * B * 4 if m["host_call"] else B * 4,
                    "path": ("ngu_epi_step_host per step, synchronous: q+alpha copied into the library's pinned buffer and pulled "
                             "over PCIe by a kernel of the (pre-armed) step, r_int+r_epi written by the epilogue into mapped host memory")
                            if m["host_call"] else "pinned H2D copies + sharded step + D2H of the rewards per step, synchronous"},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": m["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": m["achieved_gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "launch_ms": m["scan_launch_ms"],
                         "algorithmic_bytes_per_launch": bytes_per_query(N) * B,
                         "step_gbs": bytes_per_query(N) * B / (m["ms_per_step"] / 1e3) / 1e9,
                         "step_note": "step_gbs = algorithmic bytes / ms_per_step of the timed region: with pipelined steps the "
                                      "stream never drains between launches, so the whole step can beat the kernel timed alone "
                                      "(and the copy-measured peak: the scan only reads)",
                         "stream_only_launch_ms": stream_ms,
                         "stream_only_pipelined_ms_per_step": stream_piped_ms,
                         "stream_only_pipelined_gbs": (bytes_per_query(N) * B / (stream_piped_ms / 1e3) / 1e9) if stream_piped_ms else None,
                         "stream_only_note": "same kernel, distance math / top-k / flush / merge switched off "
                                             "(NGU_EPI_DEBUG_SCAN=15, results invalid, never a throughput): the TMA stream alone, "
                                             "live on this box - launched alone (launch_ms) and as pipelined steps "
                                             "(pipelined_ms_per_step = the read-only ceiling `step_gbs` should be read against)"},
        }
        if strong is not None:
            line["strong"] = strong
        if learner is not None:
            line["learner_allreduce"] = learner
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(B, N)
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)[:200]}
        emit(line)
    D.close()
    if not ok:
        log("bench.py: PARITY / HEALTH CHECK FAILED (see the parity object)")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 5],
                    help="SURVEY.md 8d configuration: 3 = 256 actors/GPU x 30k (default, headline), 2 = 64 x 5k, "
                         "1 = B=1 x 30k, 5 = 1024 actors in total over the GPUs (+ learner all-reduce)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --global-actors in total, split over the GPUs (config 3 as SURVEY.md 8d writes it)")
    ap.add_argument("--global-actors", type=int, default=256)
    ap.add_argument("--actors-per-gpu", type=int, default=None)
    ap.add_argument("--capacity", type=int, default=None)
    ap.add_argument("--scan-variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-only", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    preset = {1: (1, 30000), 2: (64, 5000), 3: (ACTORS_PER_GPU, CAPACITY), 5: (max(1024 // max(args.gpus, 1), 1), CAPACITY)}[args.config]
    if args.actors_per_gpu is None:
        args.actors_per_gpu = preset[0]
    if args.capacity is None:
        args.capacity = preset[1]
    if args.config != 3:
        args.no_secondary = args.no_secondary or args.config in (1, 2)
    if args.impl == "reference":
        claim_stdout()
        run_reference_arm(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    claim_stdout()
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
