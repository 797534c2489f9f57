"""Full-size checks at BASELINE.json's sizes (256 actors x 30 000 slots, and 1 x 30 000) through
size-independent properties plus the plain-C oracle on a subset of actors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_256x30k_properties_and_c_oracle_subset(cuda_device):
    from ngu_b200.engine import EpisodicEngine
    from oracle import c_oracle
    from oracle.episodic_oracle import EpisodicOracle
    B, N, k = 256, 30000, 10
    eng = EpisodicEngine(B, N, k=k, device=0)
    g = torch.Generator(device="cuda").manual_seed(7)
    eng._memory.normal_(generator=g)
    # plant: actor 3 gets far outliers at known slots (the farthest-k must find exactly these),
    # actor 5 gets exact duplicates of one row at many slots (tie rule: lowest slots win)
    far = [29999, 17, 12345, 4, 20000, 9999, 1, 28000, 15000, 77]
    for r, s in enumerate(far):
        eng._memory[3, s] = 100.0 + r
    dup_slots = list(range(1000, 30000, 997))
    eng._memory[5, dup_slots] = 50.0
    q = torch.randn(B, 32, device="cuda", generator=g)
    mem_before = eng._memory.clone()
    r_int, r_epi = eng.step(q, None)
    d2, idx = eng.topk()
    # property 1: planted outliers, ordered by distance (largest offset farthest)
    assert idx[3].tolist() == far[::-1]
    # property 2: ties -> lowest slots, in ascending slot order
    assert idx[5].tolist() == dup_slots[:k]
    assert len(set(d2[5].tolist())) == 1
    # property 3: values descending, slots in range, bit-exact d2 recomputed on the host for the winners
    assert (np.diff(d2, axis=1) <= 0).all() and (idx >= 0).all() and (idx < N).all()
    from oracle.episodic_oracle import d2_lane8
    memc = mem_before.cpu().numpy()
    rows = np.take_along_axis(memc, idx[:, :, None].astype(np.int64), axis=1)      # [B,k,32]
    chk = d2_lane8(rows, q.cpu().numpy()[:, None, :])
    assert np.array_equal(chk.view(np.uint32), d2.view(np.uint32))
    # property 4: append went to slot 0 for every actor, nothing else changed; cursor advanced
    assert torch.equal(eng._memory[:, 0], q)
    assert torch.equal(eng._memory[:, 1:], mem_before[:, 1:])
    assert eng.get_state()["cursor"] == 1
    # oracle (plain C scan) on every actor, 32 at a time: actors of the statically scheduled part of the
    # memory and of the dynamically scheduled part take different paths through the kernel
    assert eng.scan_geometry()["dynamic_chunk_tiles"] > 0
    qc = q.cpu().numpy()
    for lo in range(0, B, 32):
        sub = list(range(lo, lo + 32))
        mem_sub = np.ascontiguousarray(memc[sub].transpose(1, 0, 2))
        oi, od = c_oracle.scan_topk(mem_sub, qc[sub], k, True)
        assert np.array_equal(oi.T, idx[sub]) and np.array_equal(od.T.view(np.uint32), d2[sub].view(np.uint32))
    assert eng.get_state()["exchange_timeouts"] == 0
    # rewards: closed form from the engine's own top-k through the numpy oracle epilogue
    orc = EpisodicOracle(B, 16, k=k)
    want = orc.reward_from_topk(d2.T.copy(), None)
    rel = np.abs(r_int.cpu().numpy() - want) / np.abs(want)
    assert rel.max() <= 1e-5
    # property 5: reset zeroes exactly the done actors
    done = torch.zeros(B, dtype=torch.bool)
    done[[3, 200]] = True
    eng.reset(done)
    torch.cuda.synchronize()
    assert not eng._memory[3].any() and not eng._memory[200].any()
    assert torch.equal(eng._memory[4], mem_before[4].index_put((torch.tensor([0], device="cuda"),), q[4:5]))
    eng.close()


def test_idempotent_rescan_and_single_actor_30k(cuda_device):
    """Scanning twice without an append gives identical k-NN (the scan has no side effects on the memory)."""
    from ngu_b200.engine import EpisodicEngine
    from oracle import c_oracle
    B, N, k = 1, 30000, 10
    eng = EpisodicEngine(B, N, k=k, device=0)
    g = torch.Generator(device="cuda").manual_seed(11)
    eng._memory.normal_(generator=g)
    q = torch.randn(B, 32, device="cuda", generator=g)
    eng.scan(q)
    a = eng.topk()
    eng.scan(q)
    b = eng.topk()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    mem = np.ascontiguousarray(eng._memory.cpu().numpy().transpose(1, 0, 2))
    oi, od = c_oracle.scan_topk(mem, q.cpu().numpy(), k, True)
    assert np.array_equal(oi.T, a[1]) and np.array_equal(od.T.view(np.uint32), a[0].view(np.uint32))
    eng.close()


def test_long_trajectory_64x5000_vs_oracle(cuda_device):
    """BASELINE configs[1] shape: 64 actors x 5 000 slots, 25 steps with resets, against the C-scan oracle."""
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N, k = 64, 5000, 10
    rng = np.random.default_rng(5)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N, k=k, use_c_scan=True)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, device=0)
    eng.load_memory(mem0)
    for t in range(25):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        alpha = (1 + 2 * rng.standard_normal(B)).astype(np.float32)
        want = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, want["idx"])
        assert (np.abs(r_int - want["reward"]) / np.abs(want["reward"])).max() <= 1e-5
        done = rng.random(B) < 0.05
        orc.reset(done)
        eng.reset_host(done)
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


def test_256x30k_zero_memory_all_ties(cuda_device):
    """The state every real run starts in: an all-zero memory, every slot tying with every other (the
    reference's unfilled slots are real neighbours).  Ties must resolve to the lowest slots across the
    hundreds of runs the dynamic schedule cuts an actor into; three steps so that the appended rows and a
    reset take part.  Checked for every actor against the plain-C oracle."""
    from ngu_b200.engine import EpisodicEngine
    from oracle import c_oracle
    B, N, k = 256, 30000, 10
    eng = EpisodicEngine(B, N, k=k, device=0)
    assert eng.scan_geometry()["dynamic_chunk_tiles"] > 0
    g = torch.Generator(device="cuda").manual_seed(3)
    mem = np.zeros((N, B, 32), np.float32)
    for t in range(3):
        q = torch.randn(B, 32, device="cuda", generator=g)
        if t == 1:
            q[7] = 0.0                                   # distance 0 to every zero slot, > 0 to the appended one
        eng.step(q, None)
        d2, idx = eng.topk()
        qc = q.cpu().numpy()
        for lo in range(0, B, 64):
            sub = slice(lo, lo + 64)
            oi, od = c_oracle.scan_topk(np.ascontiguousarray(mem[:, sub]), qc[sub], k, True)
            assert np.array_equal(oi.T, idx[sub]), f"step {t}, actors {lo}..{lo + 63}: kNN slots differ"
            assert np.array_equal(od.T.view(np.uint32), d2[sub].view(np.uint32))
        mem[t] = qc                                      # insert_state at the shared cursor
        if t == 0:
            done = torch.zeros(B, dtype=torch.bool)
            done[[5, 100]] = True
            eng.reset(done)
            mem[:, [5, 100]] = 0.0
    assert eng.get_state()["exchange_timeouts"] == 0
    eng.close()


def test_soak_dynamic_vs_static_schedule_bit_identical(cuda_device, monkeypatch):
    """1500 steps at 256 x 30k: the dynamic schedule (chunk counter, fence-free candidate publication,
    floors, incremental merge jobs) and the static one (deterministic slots, last arriver merges) are two
    independent routes to the same k-NN, so their slots, d^2, rewards, memories and running statistics must
    agree bit for bit on every step - a soak for rare orderings (in-flight candidates, epoch tags, counter
    re-arming) that short trajectories do not hit."""
    from ngu_b200.engine import EpisodicEngine
    B, N = 256, 30000
    dyn = EpisodicEngine(B, N, device=0)
    monkeypatch.setenv("NGU_EPI_SCHED", "1.0,0")
    sta = EpisodicEngine(B, N, device=0)
    assert dyn.scan_geometry()["dynamic_chunk_tiles"] > 0 and sta.scan_geometry()["dynamic_chunk_tiles"] == 0
    g = torch.Generator(device="cuda").manual_seed(17)
    dyn._memory.normal_(generator=g)
    dyn._memory[:, 20000:] = 0.0                          # a third of every ring still unfilled: ties
    sta._memory.copy_(dyn._memory)
    for t in range(1500):
        q = torch.randn(B, 32, device="cuda", generator=g)
        a = 1.0 + 2.0 * torch.randn(B, device="cuda", generator=g)
        r1, e1 = dyn.step(q, a)
        r2, e2 = sta.step(q, a)
        if t % 50 == 0 or t > 1480:
            assert torch.equal(r1, r2) and torch.equal(e1, e2), f"step {t}: rewards differ"
            (d2a, ia), (d2b, ib) = dyn.topk(), sta.topk()
            assert np.array_equal(ia, ib) and np.array_equal(d2a.view(np.uint32), d2b.view(np.uint32)), f"step {t}"
        if t % 97 == 5:
            done = torch.rand(B, device="cuda", generator=g) < 0.05
            dyn.reset(done)
            sta.reset(done)
    assert torch.equal(dyn._memory, sta._memory)
    s1, s2 = dyn.get_state(), sta.get_state()
    assert s1["exchange_timeouts"] == 0 and s2["exchange_timeouts"] == 0
    for key in ("ed_count", "cursor", "steps", "epi_mean", "intr_mean"):
        assert s1[key] == s2[key], key
    assert np.array_equal(s1["ed_mean"], s2["ed_mean"]) and np.array_equal(s1["ed_var"], s2["ed_var"])
    dyn.close(); sta.close()


@pytest.mark.parametrize("pipelined", [False, True])
def test_back_to_back_steps_overlap_vs_oracle(cuda_device, pipelined):
    """(pipelined=True: the same through ngu_epi_step_pipelined - each scan STREAMS while the previous epilogue is
    still appending and takes the slot being appended from the previous q.)
    40 steps enqueued back to back (no host synchronisation in between) at a shape that takes the
    dynamic schedule: every scan is programmatically launched behind the previous epilogue and prefetches
    its first tiles into L2 while that epilogue is still appending - the appended rows (cursor 0..39, i.e.
    inside prefetched tiles) must nevertheless be seen by the following steps.  Every step's rewards are
    kept on the device and compared with the oracle afterwards, as are the memory and the trackers."""
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N, k, steps = 64, 30000, 10, 40
    rng = np.random.default_rng(21)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[:64] = 0.0                                       # the slots about to be overwritten start as ties
    eng = EpisodicEngine(B, N, k=k, device=0)
    assert eng.scan_geometry()["dynamic_chunk_tiles"] > 0
    eng.load_memory(mem0)
    qs = rng.standard_normal((steps, B, 32)).astype(np.float32) * 3.0     # far rows: they enter the farthest-k sets
    al = (1.0 + 2.0 * rng.standard_normal((steps, B))).astype(np.float32)
    q_dev, a_dev = torch.from_numpy(qs).cuda(), torch.from_numpy(al).cuda()
    out = torch.empty(steps, B, device="cuda")
    epi = torch.empty(steps, B, device="cuda")
    torch.cuda.synchronize()
    for t in range(steps):                                # enqueue only
        if pipelined:                                     # (nothing else may be enqueued between pipelined steps)
            eng.step(q_dev[t], a_dev[t], pipelined=True, out=(out[t], epi[t]))
        else:
            r_int, _ = eng.step(q_dev[t], a_dev[t])
            out[t].copy_(r_int)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    orc = EpisodicOracle(B, N, k=k, use_c_scan=True)
    orc.memory[...] = mem0
    for t in range(steps):
        want = orc.step(qs[t], al[t])
        rel = np.abs(got[t] - want["reward"]) / np.abs(want["reward"])
        assert rel.max() <= 1e-5, f"step {t}: reward rel err {rel.max():.2e}"
    assert np.array_equal(eng.dump_memory(), orc.memory)
    st = eng.get_state()
    np.testing.assert_allclose(st["ed_mean"], orc.ed_rms.mean, rtol=1e-6)
    assert st["exchange_timeouts"] == 0 and st["cursor"] == steps
    eng.close()


def test_256x30k_pipelined_equals_ordinary_bitwise(cuda_device):
    """The headline shape (dynamic schedule, 296 CTAs, 983 MB): 10 steps through ngu_epi_step and through
    ngu_epi_step_pipelined give identical reward bits, k-NN lists, memory and statistics."""
    from ngu_b200.engine import EpisodicEngine
    B, N, T = 256, 30000, 10
    g = torch.Generator(device="cuda").manual_seed(3)
    mem = torch.randn(B, N, 32, device="cuda", generator=g)
    qs = torch.randn(T, B, 32, device="cuda", generator=g) * 2.0
    al = 1.0 + torch.randn(T, B, device="cuda", generator=g)
    res = []
    for pipelined in (False, True):
        eng = EpisodicEngine(B, N, device=0)
        eng._memory.copy_(mem)
        r = torch.empty(T, B, device="cuda")
        e = torch.empty(T, B, device="cuda")
        torch.cuda.synchronize()
        for t in range(T):
            eng.step(qs[t], al[t], pipelined=pipelined, out=(r[t], e[t]))
        torch.cuda.synchronize()
        st = eng.get_state()
        d2, idx = eng.topk()
        res.append((r.cpu().numpy(), e.cpu().numpy(), d2, idx, st, eng._memory[:, :T + 1].cpu().numpy()))
        eng.close()
        del eng
        torch.cuda.empty_cache()
    a, b = res
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)) and np.array_equal(a[3], b[3])
    assert np.array_equal(a[5], b[5])
    assert np.array_equal(a[4]["ed_mean"], b[4]["ed_mean"]) and a[4]["cursor"] == b[4]["cursor"] == T
    assert a[4]["exchange_timeouts"] == b[4]["exchange_timeouts"] == 0
