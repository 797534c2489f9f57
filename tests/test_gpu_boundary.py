"""GPU tests of the C-ABI boundary behaviour added in round 2: partial state updates that keep the
library's bookkeeping, the sticky timeout error, stream ordering of the *_host calls behind work the
caller enqueued on its own stream, and the current-device guard."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


def test_set_state_and_set_cursor_keep_bookkeeping(cuda_device):
    """get -> modify -> set must lose nothing: the reward stash of the last fused step is folded into the
    log-only trackers, not dropped; set_cursor moves only the cursor (EpisodicNovelty.memory_idx)."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N = 6, 500
    rng = np.random.default_rng(5)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N); orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, device=0); eng.load_memory(mem0)
    for t in range(6):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        want = orc.step(q, None)
        r_int, _ = eng.step(torch.from_numpy(q).cuda(), None)      # fused step: rewards are stashed for the trackers
        if t == 2:
            eng.set_cursor(100); orc.cursor = 100                   # partial update right behind a fused step
        if t == 4:
            st = eng.get_state()
            eng.set_state({"cursor": 7, "ed_count": st["ed_count"]}); orc.cursor = 7
        assert _rel(r_int.cpu().numpy().astype(np.float64), want["reward"]).max() <= 1e-5
    st = eng.get_state()
    assert st["epi_count"] == pytest.approx(1e-4 + 6 * B)          # every step's batch reached the log-only trackers
    assert st["intr_count"] == pytest.approx(1e-4 + 6 * B)
    assert st["cursor"] == orc.cursor
    assert np.array_equal(eng.dump_memory(), orc.memory)
    with pytest.raises(Exception):
        eng.set_cursor(N)
    eng.close()


def test_sticky_timeout_error_is_loud(cuda_device):
    """A merge / exchange that gave up waiting must not produce plausible numbers with rc = 0: rewards become
    NaN, the next step call returns NGU_ESTATE, get_state raises, set_state cannot clear the counter."""
    import torch
    from ngu_b200._cabi import NguEpiError
    from ngu_b200.engine import EpisodicEngine
    B, N = 4, 300
    eng = EpisodicEngine(B, N, device=0)
    q = torch.randn(B, 32, device="cuda")
    r, _ = eng.step(q, None)
    assert torch.isfinite(r).all()
    assert eng._lib.ngu_epi_debug_inject_timeout(eng._h) == 0
    r, r_epi = eng.step(q, None)            # launches (the host cannot know yet), but the kernel poisons its output
    torch.cuda.synchronize()
    assert torch.isnan(r).all() and torch.isnan(r_epi).all()
    with pytest.raises(NguEpiError, match="timed out"):
        eng.step(q, None)
    with pytest.raises(NguEpiError, match="timed out"):
        eng.step_host(np.zeros((B, 32), np.float32))
    with pytest.raises(NguEpiError, match="timed out"):
        eng.get_state()
    st = eng.get_state(allow_failed=True)
    assert st["exchange_timeouts"] == 1
    eng.set_state({"cursor": 0})
    assert eng.get_state(allow_failed=True)["exchange_timeouts"] == 1
    eng.close()


def test_host_calls_are_ordered_after_user_stream_work(cuda_device):
    """insert_state / device-mask reset on the caller's stream, immediately followed by reset_host / step_host on the
    library's private stream: the private stream must wait (an appended row could land after the
    zero-fill, a scan could read a half-zeroed memory)."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    B, N = 48, 20000
    rng = np.random.default_rng(11)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N, use_c_scan=True); orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, device=0); eng.load_memory(mem0)
    side = torch.cuda.Stream()
    for t in range(6):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        done_dev = rng.random(B) < 0.5
        done_host = rng.random(B) < 0.3
        q_ins = rng.standard_normal((B, 32), dtype=np.float32)
        with torch.cuda.stream(side):
            # a long-running filler first, so that the work below is still queued when the *_host call arrives
            filler = torch.empty(1 << 28, dtype=torch.uint8, device="cuda"); filler.zero_(); filler.zero_()
            eng.reset(torch.from_numpy(done_dev).cuda())
            eng.append(torch.from_numpy(q_ins).cuda())
        orc.reset(done_dev); orc.append(q_ins)
        eng.reset_host(done_host); orc.reset(done_host)
        want = orc.step(q, None)
        r_int, _ = eng.step_host(q)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, want["idx"]), f"step {t}"
        assert _rel(r_int.astype(np.float64), want["reward"]).max() <= 1e-5
        del filler
    assert np.array_equal(eng.dump_memory(), orc.memory)
    eng.close()


def test_calls_restore_the_current_device(cuda_device):
    """Every entry point runs on the handle's GPU and puts the caller's current device back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ngu_b200.engine import EpisodicEngine
    torch.cuda.set_device(0)
    eng = EpisodicEngine(4, 300, device=1)
    assert torch.cuda.current_device() == 0
    r, _ = eng.step_host(np.random.default_rng(0).standard_normal((4, 32), dtype=np.float32))
    dev = C.c_int(-1)
    C.CDLL("libcudart.so.12").cudaGetDevice(C.byref(dev))
    assert dev.value == 0 and np.isfinite(r).all()
    eng.get_state(); eng.topk()
    C.CDLL("libcudart.so.12").cudaGetDevice(C.byref(dev))
    assert dev.value == 0
    eng.close()


@pytest.mark.parametrize("B,N", [(5, 700), (64, 5000), (256, 2000)])
def test_pre_armed_host_steps_vs_oracle(cuda_device, monkeypatch, B, N):
    """ngu_epi_step_host in a tight loop: from the second call on the next step's launches are enqueued ahead and
    released by the doorbell (armed_stage_kernel).  Every way such a step can end is exercised against the oracle:
    doorbell in time (with and without modulators), the window running out while the caller sleeps, cancellation by
    another entry point (topk / reset_host / get_state / a device-pointer step)."""
    import time
    import torch
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    monkeypatch.setenv("NGU_EPI_HOST_PROFILE", "1")
    monkeypatch.setenv("NGU_EPI_ARM_WINDOW_US", "300")
    rng = np.random.default_rng(B * 7 + N)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    orc = EpisodicOracle(B, N, use_c_scan=True); orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, device=0); eng.load_memory(mem0)
    T = 60
    qs = rng.standard_normal((T, B, 32), dtype=np.float32)
    al = (1.0 + 2.0 * rng.standard_normal((T, B))).astype(np.float32)
    got = []
    t = 0
    while t < T:
        kind = (t // 6) % 5
        if kind in (0, 1):                       # tight loop: armed steps run
            for _ in range(6):
                got.append(eng.step_host(qs[t], al[t] if (t % 2 or kind == 0) else None)[0]); t += 1
        elif kind == 2:                          # the caller goes away for longer than the window
            for _ in range(6):
                got.append(eng.step_host(qs[t], al[t])[0]); t += 1
                if t % 2:
                    time.sleep(0.002)
        elif kind == 3:                          # other entry points in between cancel the armed step
            for i in range(6):
                got.append(eng.step_host(qs[t], al[t])[0]); t += 1
                if i % 3 == 0:
                    eng.topk()
                elif i % 3 == 1:
                    eng.get_state()
                else:
                    eng.reset_host(np.zeros(B, np.uint8) + (np.arange(B) == (t % B)))
                    orc_done = (np.arange(B) == (t % B))
                    got.append(("reset", orc_done))
        else:                                    # device-pointer steps interleaved with host steps
            for i in range(6):
                if i % 2:
                    r = eng.step(torch.from_numpy(qs[t]).cuda(), torch.from_numpy(al[t]).cuda())[0].cpu().numpy()
                else:
                    r = eng.step_host(qs[t], al[t])[0]
                got.append(r); t += 1
    # replay on the oracle
    t = 0
    worst = 0.0
    for item in got:
        if isinstance(item, tuple):
            orc.reset(item[1]); continue
        kind = (t // 6) % 5
        alpha = al[t] if not (kind == 1 and t % 2 == 0) else None
        want = orc.step(qs[t], alpha)["reward"]
        worst = max(worst, float(_rel(item.astype(np.float64), want).max()))
        t += 1
    assert worst <= 1e-5, worst
    assert np.array_equal(eng.dump_memory(), orc.memory)
    st = eng.get_state()
    assert st["exchange_timeouts"] == 0 and st["steps"] == T
    eng.close()


def test_step_host_writes_into_the_callers_buffer(cuda_device):
    """step_host(out=...): the rewards land in the caller's (2, B) array, no intermediate copy; same values."""
    from ngu_b200.engine import EpisodicEngine
    B, N = 9, 800
    rng = np.random.default_rng(2)
    mem0 = rng.standard_normal((N, B, 32), dtype=np.float32)
    e1 = EpisodicEngine(B, N, device=0); e1.load_memory(mem0)
    e2 = EpisodicEngine(B, N, device=0); e2.load_memory(mem0)
    out = np.zeros((2, B), np.float32)
    for t in range(6):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        a = (1 + rng.standard_normal(B)).astype(np.float32)
        r1, p1 = e1.step_host(q, a)
        r2, p2 = e2.step_host(q, a, out=out)
        assert r2.base is out and np.array_equal(r1, out[0]) and np.array_equal(p1, out[1])
    with pytest.raises(ValueError):
        e2.step_host(q, a, out=np.zeros((B, 2), np.float32))
    e1.close(); e2.close()
