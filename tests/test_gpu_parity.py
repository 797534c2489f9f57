"""GPU parity: the CUDA path (through the C ABI) against the golden trajectories
recorded from the unmodified reference and against the numpy oracle.

Bars (BASELINE.json north_star): kNN slot indices bit-exact (ties -> lowest
slot), d^2 bit-exact, rewards within 1e-5 relative in fp32.
"""
import numpy as np
import pytest

from tests.conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5   # north_star: "intrinsic rewards within 1e-5 relative in fp32"


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("variant", [0, 2])
def test_golden_trajectory(cuda_device, name, variant):
    from ngu_b200.engine import EpisodicEngine
    g = load_golden(name)
    eng = EpisodicEngine(g["B"], g["N"], k=g["k"], device=0, scan_variant=variant)
    if g["mem0_full"].any():
        eng.load_memory(g["mem0_full"])
    for t in range(g["steps"]):
        r_int, r_epi = eng.step_host(g["q"][t], g["alpha"][t])
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, g["idx"][t]), f"step {t}: kNN slots differ"
        assert np.array_equal(d2.T.view(np.uint32), g["d2"][t].view(np.uint32)), f"step {t}: d^2 bits differ"
        rel = _rel(r_int.astype(np.float64), g["reward"][t])
        assert rel.max() <= REL_TOL, f"step {t}: reward rel err {rel.max():.3e}"
        st = eng.get_state()
        np.testing.assert_allclose(st["ed_mean"], g["mean"][t], rtol=1e-6)
        np.testing.assert_allclose(st["ed_var"], g["var"][t], rtol=1e-4, atol=1e-9)
        assert st["ed_count"] == g["count"][t]
        # log-only trackers against the reference's own (f64 batch sums here, f32 / f64 numpy sums there)
        np.testing.assert_allclose([st["epi_mean"], st["epi_var"], st["epi_count"]], g["epi_rms"][t], rtol=1e-5)
        np.testing.assert_allclose([st["intr_mean"], st["intr_var"], st["intr_count"]], g["intr_rms"][t], rtol=1e-5)
        eng.reset_host(g["done"][t])
    st = eng.get_state()
    assert st["cursor"] == int(g["cursor_final"])
    mem = eng.dump_memory()
    if g["mem_final"].size:
        assert np.array_equal(mem, g["mem_final"])
    else:
        assert abs(float(mem.astype(np.float64).sum()) - float(g["mem_final_sum"])) <= 1e-6 * max(1.0, abs(float(g["mem_final_sum"])))
    eng.close()


@pytest.mark.parametrize("B,N,k,mode", [(16, 3000, 10, "reference"), (3, 1000, 1, "reference"),
                                         (5, 2049, 32, "reference"), (9, 4097, 10, "paper"),
                                         (2, 33, 10, "reference"), (1, 30000, 10, "reference"),
                                         # many actors: the pairwise rank sums have 8 / 64 leaves per column
                                         (1024, 100, 10, "reference"), (5000, 40, 4, "reference")])
def test_random_vs_oracle(cuda_device, B, N, k, mode):
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    rng = np.random.default_rng(B * 1000 + N + k)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[rng.random(N) < 0.2] = 0.0                       # ties among zero slots
    orc = EpisodicOracle(B, N, k=k, mode=mode)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, mode=mode, device=0)
    eng.load_memory(mem0)
    for t in range(6):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)
        ref = orc.step(q, alpha)
        r_int, r_epi = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"])
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32))
        assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
        done = rng.random(B) < 0.3
        orc.reset(done)
        eng.reset_host(done)
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


def test_adversarial_ascending_distances(cuda_device):
    """Every slot beats the running threshold (distances ascending in slot order):
    exercises the bitonic batch-merge path of the warp top-k."""
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N, k = 3, 5000, 10
    mem0 = np.zeros((N, B, 32), np.float32)
    mem0[:, :, 0] = np.arange(N, dtype=np.float32)[:, None] * 0.01
    mem0[:, 1, 0] = mem0[::-1, 1, 0]                      # descending for actor 1
    orc = EpisodicOracle(B, N, k=k); orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, device=0); eng.load_memory(mem0)
    q = np.zeros((B, 32), np.float32)
    ref = orc.step(q, None)
    r_int, _ = eng.step_host(q, None)
    d2, idx = eng.topk()
    assert np.array_equal(idx.T, ref["idx"])
    assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
    eng.close()


def test_device_pointer_path_matches_host_path(cuda_device):
    """ngu_epi_step (device pointers, caller's stream) == ngu_epi_step_host."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    B, N = 8, 2000
    rng = np.random.default_rng(7)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    e1 = EpisodicEngine(B, N, device=0); e1.load_memory(mem0)
    e2 = EpisodicEngine(B, N, device=0); e2.load_memory(mem0)
    for t in range(5):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        a = (1 + rng.standard_normal(B)).astype(np.float32)
        r1, _ = e1.step_host(q, a)
        r2, _ = e2.step(torch.from_numpy(q).cuda(), torch.from_numpy(a).cuda())
        assert np.array_equal(r1, r2.cpu().numpy())
    assert torch.equal(e1.episodic_memory, e2.episodic_memory)
    assert e1.episodic_memory.shape == (N, B, 32)
    e1.close(); e2.close()


def test_create_errors(cuda_device):
    from ngu_b200.engine import EpisodicEngine
    from ngu_b200._cabi import NguEpiError
    with pytest.raises(NguEpiError):
        EpisodicEngine(4, 5, k=10, device=0)       # k > capacity: the reference's topk raises too
    with pytest.raises(NguEpiError):
        EpisodicEngine(4, 64, k=33, device=0)


# ---- dynamic work schedule of the scan kernel (chunk counter, fence-free candidate publication,
# ---- floors, incremental merge jobs) -------------------------------------------------------------
def _check_partition(eng, B, N):
    geo = eng.scan_geometry()
    tpa = -(-N // geo["tile_rows"])
    cover = np.zeros(B * tpa, np.int32)
    chunks = eng.schedule_chunks()
    assert len(chunks) == geo["chunks"]
    for t, te in chunks:
        assert 0 <= t <= te <= B * tpa
        cover[t:te] += 1
    assert (cover == 1).all(), "the schedule must hand out every tile exactly once"
    return geo


@pytest.mark.parametrize("B,N,sched", [(256, 30000, None), (37, 30000, None), (1, 30000, None), (64, 5000, None),
                                        (300, 7777, "0.3,8,1"), (24, 12000, "0.5,2,1"), (24, 12000, "0.0,3,1")])
def test_schedule_partitions_the_tile_space(cuda_device, monkeypatch, B, N, sched):
    from ngu_b200.engine import EpisodicEngine
    if sched:
        monkeypatch.setenv("NGU_EPI_SCHED", sched)
    eng = EpisodicEngine(B, N, device=0)
    geo = _check_partition(eng, B, N)
    if sched or (B, N) == (256, 30000):
        assert geo["dynamic_chunk_tiles"] > 0, "expected the dynamic schedule for this shape"
    eng.close()


@pytest.mark.parametrize("sched,epoch,k,mode,poll", [
    ("0.5,2,1", None, 10, "reference", None), ("0.0,3,1", None, 10, "reference", None),
    ("0.6,16,1", None, 10, "reference", None), ("0.5,2,1", "0xFFFFFFFD", 10, "reference", None),
    ("0.5,2,1", None, 32, "reference", None), ("0.3,4,1", None, 1, "reference", None),
    ("0.5,2,1", None, 10, "paper", None), ("0.5,2,1", None, 10, "reference", "0"), ("0.6,16,1", None, 10, "reference", "0")])
def test_dynamic_schedule_vs_oracle(cuda_device, monkeypatch, sched, epoch, k, mode, poll):
    """Many short runs per actor (hundreds of flushes, floors, in-flight candidates, merge jobs) at a size
    the oracle checks in full; `epoch` starts the launch epoch just below its 32-bit wrap; `poll` = "0"
    makes every merge job that is not complete at its first look give up, so that the last warp of the
    launch merges the leftovers (the path that keeps a not fully co-resident grid from deadlocking)."""
    if poll is not None:
        monkeypatch.setenv("NGU_EPI_DEBUG_POLL_LIMIT", poll)
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    monkeypatch.setenv("NGU_EPI_SCHED", sched)
    if epoch:
        monkeypatch.setenv("NGU_EPI_DEBUG_EPOCH", epoch)
    B, N = 24, 12000
    rng = np.random.default_rng(99)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[rng.random(N) < 0.2] = 0.0                       # ties among zero slots
    mem0[:, 3, :] = 0.0                                   # an actor whose slots all tie
    mem0[:, 4, 1:] = 0.0
    mem0[:, 4, 0] = np.arange(N, dtype=np.float32) * 0.01  # ascending distances: every run publishes a full list
    mem0[:, 5, 1:] = 0.0
    mem0[:, 5, 0] = np.arange(N, dtype=np.float32)[::-1] * 0.01
    orc = EpisodicOracle(B, N, k=k, mode=mode, use_c_scan=True)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, mode=mode, device=0)
    assert eng.scan_geometry()["dynamic_chunk_tiles"] > 0
    eng.load_memory(mem0)
    for t in range(8):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        q[4] = 0.0
        q[5] = 0.0
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)
        ref = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"]), f"step {t}: kNN slots differ"
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32)), f"step {t}: d^2 bits differ"
        assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
        done = rng.random(B) < 0.1
        done[3:6] = False
        orc.reset(done)
        eng.reset_host(done)
    assert eng.get_state()["exchange_timeouts"] == 0, "a merge gave up waiting for a candidate"
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


def _same_state(a, b):
    assert a.keys() == b.keys()
    for key in a:
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key


@pytest.mark.parametrize("B,N", [(7, 600), (256, 2000), (700, 300)])
def test_deferred_log_trackers(cuda_device, B, N):
    """epi_novel_rms / intrinsic_novel_rms (episodic_novelty.py:82, intrinsic_novelty.py:32) are folded into
    the device trackers one launch late, by the next step's epilogue while its scan drains, or by
    ngu_epi_get_state.  Both routes must give the same bits, and the trackers must follow the
    reference's RunningMeanStd fed with the rewards the steps returned."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import RunningMeanStdOracle
    rng = np.random.default_rng(B + N)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    lazy = EpisodicEngine(B, N, device=0); lazy.load_memory(mem0)      # absorbs inside the epilogue kernel
    eager = EpisodicEngine(B, N, device=0); eager.load_memory(mem0)    # get_state after every step
    epi, intr = RunningMeanStdOracle(1e-4), RunningMeanStdOracle(1e-4)
    T = 25
    for t in range(T):
        q = torch.from_numpy(rng.standard_normal((B, 32), dtype=np.float32)).cuda()
        alpha = torch.from_numpy((1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)).cuda()
        r_int, r_epi = lazy.step(q, alpha)
        r2, _ = eager.step(q, alpha)
        assert torch.equal(r_int, r2)
        epi.update(r_epi.cpu().numpy()); intr.update(r_int.cpu().numpy())
        se = eager.get_state()
        if t % 6 == 5:                       # a state read in the middle of the lazy run must not double count
            _same_state(lazy.get_state(), se)
    sl = lazy.get_state()
    _same_state(sl, eager.get_state())
    assert sl["epi_count"] == epi.count and sl["intr_count"] == intr.count
    np.testing.assert_allclose(sl["epi_mean"], epi.mean, rtol=1e-6)
    np.testing.assert_allclose(sl["intr_mean"], intr.mean, rtol=1e-6)
    np.testing.assert_allclose(sl["epi_var"], epi.var, rtol=1e-3)
    np.testing.assert_allclose(sl["intr_var"], intr.var, rtol=1e-3)
    lazy.close(); eager.close()


@pytest.mark.parametrize("stage", ["kernel", "copy"])
@pytest.mark.parametrize("B", [1, 3, 64])
def test_host_step_input_staging(cuda_device, monkeypatch, stage, B):
    """ngu_epi_step_host moves the step's inputs host -> device either with a PDL-chained kernel that reads
    the pinned buffer over PCIe (large shapes by default) or with the copy engine; both against the oracle,
    including actor counts whose modulator block is not a whole number of 16-byte pieces."""
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    monkeypatch.setenv("NGU_EPI_HOST_STAGE", stage)
    N = 1200
    rng = np.random.default_rng(17 * B)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N); orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, device=0); eng.load_memory(mem0)
    for t in range(5):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32) if t % 2 == 0 else None
        ref = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"])
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32))
        assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


@pytest.mark.parametrize("B,N,k,mode", [(16, 3000, 10, "reference"), (5, 2049, 32, "reference"), (64, 5000, 10, "reference"),
                                         (1, 30000, 10, "reference"), (1024, 100, 10, "reference"), (3, 64, 10, "reference"),
                                         (9, 4097, 10, "paper"), (2, 33, 10, "reference"), (40, 96, 4, "reference")])
@pytest.mark.parametrize("cta_merge", ["0", "1"])
def test_static_merge_levels_vs_oracle(cuda_device, monkeypatch, B, N, k, mode, cta_merge):
    """Static schedules merge an actor's runs in one level (every warp publishes its list, the last arrival merges
    them) or in two (CTA combine in shared memory first, scan_topk.cuh::static_cta_finish); the library picks by the
    number of runs per actor.  Both forced here on shapes where a CTA holds part of an actor, exactly one actor,
    several whole actors and the boundaries between them; ordinary steps and a burst of pipelined ones."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    monkeypatch.setenv("NGU_EPI_CTA_MERGE", cta_merge)
    rng = np.random.default_rng(B * 31 + N + k)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[rng.random(N) < 0.2] = 0.0
    orc = EpisodicOracle(B, N, k=k, mode=mode, use_c_scan=B * N > 50000)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, mode=mode, device=0)
    assert eng.scan_geometry()["dynamic_chunk_tiles"] == 0
    eng.load_memory(mem0)
    for t in range(5):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)
        ref = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"]), f"step {t}: kNN slots differ"
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32)), f"step {t}: d^2 bits differ"
        assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
        done = rng.random(B) < 0.3
        orc.reset(done)
        eng.reset_host(done)
    T = 12
    qs = rng.standard_normal((T, B, 32), dtype=np.float32)
    al = (1.0 + 2.0 * rng.standard_normal((T, B))).astype(np.float32)
    qd, ad = torch.from_numpy(qs).cuda(), torch.from_numpy(al).cuda()
    r = torch.empty((T, B), dtype=torch.float32, device="cuda")
    e = torch.empty_like(r)
    for t in range(T):
        eng.step(qd[t], ad[t], pipelined=True, out=(r[t], e[t]))
    got = r.cpu().numpy()
    for t in range(T):
        ref = orc.step(qs[t], al[t])
        assert _rel(got[t].astype(np.float64), ref["reward"]).max() <= REL_TOL, f"pipelined step {t}"
    d2, idx = eng.topk()
    assert np.array_equal(idx.T, ref["idx"]) and np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32))
    assert np.array_equal(eng.dump_memory(), orc.memory)
    assert eng.get_state()["exchange_timeouts"] == 0
    eng.close()


@pytest.mark.parametrize("B,N", [(7, 2500), (40, 9000)])
def test_debug_capable_kernel_flavour_vs_oracle(cuda_device, monkeypatch, B, N):
    """The default scan geometry exists in two builds: the production one (experiment switches and timeline stamps
    compiled out of the streaming loop) and the debug-capable one the library launches when NGU_EPI_DEBUG_TIMELINE /
    NGU_EPI_DEBUG_SCAN is set.  Every other test runs the first; this one runs the second against the oracle and reads
    the per-warp stamps it records."""
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    monkeypatch.setenv("NGU_EPI_DEBUG_TIMELINE", "1")
    rng = np.random.default_rng(B + N)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, device=0)
    eng.load_memory(mem0)
    for t in range(4):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32)
        ref = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"])
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32))
        assert _rel(r_int.astype(np.float64), ref["reward"]).max() <= REL_TOL
    tl = np.zeros((148 * 4 * 32, 8), np.uint64)
    nw = eng._lib.ngu_epi_debug_timeline(eng._h, tl.ctypes.data, tl.shape[0])
    assert nw > 0 and (tl[:nw, 0] > 0).any() and (tl[:nw, 3] >= tl[:nw, 0]).all()   # start <= end stamps of every warp
    eng.close()
