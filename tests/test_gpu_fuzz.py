"""Randomised parity sweep: shapes, k, modes, schedule knobs and memory contents drawn from a fixed seed,
every step compared with the oracle (slots and d^2 bit-exact, rewards within 1e-5 relative, memory equal).
The fixed cases elsewhere pin the known edge cases; this one looks for the ones nobody thought of."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _draw(rng):
    kind = rng.integers(0, 6)
    if kind == 0:      # tiny memories, N around the tile size
        B, N = int(rng.integers(1, 40)), int(rng.integers(1, 70))
    elif kind == 1:    # many actors, short memories
        B, N = int(rng.integers(200, 1500)), int(rng.integers(10, 120))
    elif kind == 2:    # few actors, long memories
        B, N = int(rng.integers(1, 6)), int(rng.integers(3000, 40000))
    elif kind == 3:    # N one off a tile multiple
        B, N = int(rng.integers(2, 30)), int(32 * rng.integers(1, 200) + rng.choice([-1, 0, 1]))
    else:
        B, N = int(rng.integers(1, 120)), int(rng.integers(33, 4000))
    k = int(rng.choice([1, 2, 5, 10, 10, 10, 17, 32]))
    k = max(1, min(k, N))
    while B * (k + 2) > 40000:
        B //= 2
    mode = "paper" if rng.random() < 0.25 else "reference"
    fill = rng.choice(["randn", "zeros", "half_zero", "dups", "tiny", "huge"])
    sched = None
    if rng.random() < 0.5 and B * N >= 4000:
        sched = f"{rng.choice([0.0, 0.3, 0.5, 0.6, 0.8])},{int(rng.choice([1, 2, 3, 4, 8, 16]))},1,{int(rng.choice([1, 2, 4]))}"
    taper = f"{rng.choice([0.0, 0.2, 0.5])},{rng.choice([0.0, 0.2, 0.4])}" if sched and rng.random() < 0.5 else None
    return B, N, k, mode, fill, sched, taper


def _memory(rng, fill, N, B):
    if fill == "zeros":
        return np.zeros((N, B, 32), np.float32)
    if fill == "dups":
        base = rng.standard_normal((4, B, 32), dtype=np.float32)
        return base[rng.integers(0, 4, size=N)].copy()
    m = rng.standard_normal((N, B, 32), dtype=np.float32)
    if fill == "half_zero":
        m[rng.random(N) < 0.5] = 0.0
    if fill == "tiny":
        m *= np.float32(1e-18)       # squares underflow to denormals / zero
    if fill == "huge":
        m *= np.float32(1e15)        # squares near the top of the f32 range
    return m


@pytest.mark.parametrize("seed", range(int(os.environ.get("NGU_FUZZ_SEEDS", "48"))))   # widen the hunt with NGU_FUZZ_SEEDS=N
def test_random_configuration_vs_oracle(cuda_device, monkeypatch, seed):
    from ngu_b200.engine import EpisodicEngine
    from oracle.episodic_oracle import EpisodicOracle
    rng = np.random.default_rng(1000 + seed)
    B, N, k, mode, fill, sched, taper = _draw(rng)
    for name, val in (("NGU_EPI_SCHED", sched), ("NGU_EPI_TAPER", taper)):
        if val:
            monkeypatch.setenv(name, val)
        else:
            monkeypatch.delenv(name, raising=False)
    what = f"seed {seed}: B={B} N={N} k={k} {mode} {fill} sched={sched} taper={taper}"
    mem0 = _memory(rng, fill, N, B)
    scale = {"tiny": 1e-18, "huge": 1e15}.get(fill, 1.0)
    orc = EpisodicOracle(B, N, k=k, mode=mode, use_c_scan=B * N > 200000)
    orc.memory[...] = mem0
    eng = EpisodicEngine(B, N, k=k, mode=mode, device=0)
    eng.load_memory(mem0)
    for t in range(4):
        q = (rng.standard_normal((B, 32), dtype=np.float32) * np.float32(scale)).astype(np.float32)
        alpha = (1.0 + 2.0 * rng.standard_normal(B)).astype(np.float32) if t % 2 else None
        with np.errstate(all="ignore"):
            ref = orc.step(q, alpha)
        r_int, _ = eng.step_host(q, alpha)
        d2, idx = eng.topk()
        assert np.array_equal(idx.T, ref["idx"]), f"{what}, step {t}: kNN slots differ"
        assert np.array_equal(d2.T.view(np.uint32), ref["d2"].view(np.uint32)), f"{what}, step {t}: d^2 bits differ"
        want = np.asarray(ref["reward"], np.float64)
        got = r_int.astype(np.float64)
        fin = np.isfinite(want)
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}, step {t}: NaN pattern differs"
        assert np.array_equal(got[np.isinf(want)], want[np.isinf(want)]), f"{what}, step {t}: infinities differ"
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-30)
        assert rel.size == 0 or rel.max() <= 1e-5, f"{what}, step {t}: reward rel err {rel.max():.3e}"
        done = rng.random(B) < 0.25
        orc.reset(done)
        eng.reset_host(done)
    assert np.array_equal(eng.dump_memory(), orc.memory), what
    assert eng.get_state()["exchange_timeouts"] == 0, what
    eng.close()
