"""GPU parity of the software-pipelined step (ngu_epi_step_pipelined, include/ngu_epi.h): steps enqueued back to
back, each scan overlapping the previous step's merge tail / epilogue and taking the content of the slot that step
is appending from its q.  Results must be those of the ordinary step, bit for bit, and match the oracle
(episodic_novelty.py:37-89: query -> running-mean update -> reward -> append, one shared ring cursor).

Bars: kNN slots and d^2 bit-exact (checked on the last step of every burst, where the pipeline is deepest),
rewards <= 1e-5 relative on every step, final memory and cursor identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


def _run_burst(eng, orc, rng, B, steps, torch, check_topk=True):
    """`steps` pipelined steps enqueued without any host synchronisation, then everything is checked."""
    qs = torch.from_numpy(rng.standard_normal((steps, B, 32), dtype=np.float32)).cuda()
    al = torch.from_numpy((1.0 + 2.0 * rng.standard_normal((steps, B))).astype(np.float32)).cuda()
    r_int = torch.empty((steps, B), dtype=torch.float32, device="cuda")
    r_epi = torch.empty((steps, B), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for t in range(steps):
        eng.step(qs[t], al[t], pipelined=True, out=(r_int[t], r_epi[t]))
    torch.cuda.synchronize()
    qh, ah, got = qs.cpu().numpy(), al.cpu().numpy(), r_int.cpu().numpy()
    ref = None
    for t in range(steps):
        ref = orc.step(qh[t], ah[t])
        rel = _rel(got[t].astype(np.float64), ref["reward"])
        assert rel.max() <= REL_TOL, f"step {t} of the burst: reward rel err {rel.max():.3e}"
    if check_topk:
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"]), "kNN slots of the burst's last step differ"
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32)), "d^2 bits differ"


@pytest.mark.parametrize("B,N,k,sched", [
    (8, 2500, 10, None),            # static schedule, a few CTAs: several scans co-resident
    (1, 30000, 10, None),           # one actor cut into 235 runs
    (3, 64, 10, None),              # ring wraps many times: the appended slot walks through every tile
    (2, 10, 10, None),              # N == k
    (5, 2049, 32, None),            # ragged last tile, k = 32
    (64, 5000, 10, None),           # config 2
    (48, 6000, 10, "0.5,4,1,1"),    # dynamic schedule forced on a small shape: reservations, floors, merge jobs
    (48, 6000, 10, "0.0,8,1,1"),    # no static share at all
])
def test_pipelined_bursts_vs_oracle(cuda_device, monkeypatch, B, N, k, sched):
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    if sched:
        monkeypatch.setenv("NGU_EPI_SCHED", sched)
    rng = np.random.default_rng(B * 7919 + N + k)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    mem0[rng.random(N) < 0.2] = 0.0
    orc = EpisodicOracle(B, N, k=k)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, device=0)
    eng.load_memory(mem0)
    for burst, steps in enumerate((1, 2, 7, 40 if N <= 5000 else 12)):
        _run_burst(eng, orc, rng, B, steps, torch)
        st = eng.get_state()                      # closes the chain; the next burst starts with an ordinary launch
        assert st["cursor"] == orc.cursor and st["exchange_timeouts"] == 0
        np.testing.assert_allclose(st["ed_mean"], orc.ed_rms.mean, rtol=1e-6)
        if burst == 1:                            # a reset between bursts (intrinsic_novelty.py:36-40)
            done = rng.random(B) < 0.4
            orc.reset(done)
            eng.reset_host(done)
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


def test_pipelined_equals_ordinary_bitwise(cuda_device):
    """Same inputs through ngu_epi_step and ngu_epi_step_pipelined: identical reward bits, memory, statistics."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    B, N, T = 37, 3000, 60
    rng = np.random.default_rng(11)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    qs = torch.from_numpy(rng.standard_normal((T, B, 32), dtype=np.float32)).cuda()
    al = torch.from_numpy((1 + rng.standard_normal((T, B))).astype(np.float32)).cuda()
    outs = []
    for pipelined in (False, True):
        eng = EpisodicEngine(B, N, device=0)
        eng.load_memory(mem0)
        r = torch.empty((T, B), dtype=torch.float32, device="cuda")
        e = torch.empty((T, B), dtype=torch.float32, device="cuda")
        for t in range(T):
            eng.step(qs[t], al[t], pipelined=pipelined, out=(r[t], e[t]))
        torch.cuda.synchronize()
        st = eng.get_state()
        outs.append((r.cpu().numpy(), e.cpu().numpy(), eng.dump_memory(), st, eng.topk()))
        eng.close()
    a, b = outs
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[4][0].view(np.uint32), b[4][0].view(np.uint32)) and np.array_equal(a[4][1], b[4][1])
    for key in ("ed_mean", "ed_var"):
        assert np.array_equal(a[3][key], b[3][key])
    for key in ("ed_count", "cursor", "steps", "epi_mean", "epi_var", "epi_count", "intr_mean", "intr_var", "intr_count"):
        assert a[3][key] == b[3][key], key


def test_pipelined_mixed_with_other_calls(cuda_device):
    """Anything else on the handle ends the chain: the next pipelined call must behave like an ordinary step
    (appends through insert_state, device-mask resets, cursor moves, host steps in between)."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N, k = 6, 700, 10
    rng = np.random.default_rng(5)
    orc = EpisodicOracle(B, N, k=k)
    eng = EpisodicEngine(B, N, k=k, device=0)
    for t in range(60):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        a = (1 + rng.standard_normal(B)).astype(np.float32)
        ref = orc.step(q, a)
        kind = t % 6
        if kind == 4:
            got, _ = eng.step_host(q, a)
        else:
            r, _ = eng.step(torch.from_numpy(q).cuda(), torch.from_numpy(a).cuda(), pipelined=kind != 5)
            got = r.cpu().numpy()
        assert _rel(got.astype(np.float64), ref["reward"]).max() <= REL_TOL, f"step {t}"
        if kind == 1:
            extra = rng.standard_normal((B, 32), dtype=np.float32)
            orc.append(extra)
            eng.append(torch.from_numpy(extra).cuda())
        elif kind == 2:
            done = rng.random(B) < 0.3
            orc.reset(done)
            eng.reset(torch.from_numpy(done).cuda())
        elif kind == 3:
            c = int(rng.integers(0, N))
            orc.cursor = c
            eng.set_cursor(c)
    assert np.array_equal(eng.dump_memory(), orc.memory)
    assert eng.get_state()["cursor"] == orc.cursor
    eng.close()


def test_pipelined_long_run_256x300(cuda_device):
    """Many actors, ring wraps twice, 700 steps in one burst (rank sums with several pairwise leaves)."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N = 256, 300
    rng = np.random.default_rng(3)
    orc = EpisodicOracle(B, N)
    eng = EpisodicEngine(B, N, device=0)
    _run_burst(eng, orc, rng, B, 700, torch)
    assert np.array_equal(eng.dump_memory(), orc.memory)
    assert eng.get_state()["exchange_timeouts"] == 0
    eng.close()


def test_pipelined_with_inputs_that_need_conversion_falls_back(cuda_device):
    """Inputs the engine has to convert (float64, non-contiguous) are produced by a copy enqueued between the steps:
    the pipelined contract does not hold, so the engine launches those steps the ordinary way - same results."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N, T = 12, 1500, 20
    rng = np.random.default_rng(8)
    orc = EpisodicOracle(B, N)
    eng = EpisodicEngine(B, N, device=0)
    qs = rng.standard_normal((T, B, 32), dtype=np.float32)
    al = (1 + rng.standard_normal((T, B))).astype(np.float32)
    q64 = torch.from_numpy(qs.astype(np.float64)).cuda()                 # dtype conversion needed
    wide = torch.zeros(T, B, 64, device="cuda")
    wide[:, :, ::2] = torch.from_numpy(qs).cuda()                        # non-contiguous view
    a_dev = torch.from_numpy(al).cuda()
    q32 = torch.from_numpy(qs).cuda()
    r = torch.empty(T, B, device="cuda")
    e = torch.empty(T, B, device="cuda")
    torch.cuda.synchronize()
    for t in range(T):
        q = q64[t] if t % 3 == 0 else (wide[t, :, ::2] if t % 3 == 1 else q32[t])      # every third step is truly pipelined
        eng.step(q, a_dev[t], pipelined=True, out=(r[t], e[t]))
    torch.cuda.synchronize()
    got = r.cpu().numpy()
    for t in range(T):
        want = orc.step(qs[t], al[t])
        assert _rel(got[t].astype(np.float64), want["reward"]).max() <= REL_TOL, f"step {t}"
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()
