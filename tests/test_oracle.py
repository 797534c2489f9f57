"""CPU tests: the oracle (numpy, plain C, torch port) against the golden trajectories
recorded from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from tests.conftest import golden_names, load_golden

REL_TOL = 1e-5
SMALL = [n for n in golden_names() if n not in ("b64_n5000",)]


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("use_c", [False, True])
def test_oracle_matches_reference_golden(name, use_c):
    from oracle.episodic_oracle import EpisodicOracle
    g = load_golden(name)
    if not use_c and g["B"] * g["N"] > 100000:
        pytest.skip("numpy scan too slow at this size; covered by the C scan")
    orc = EpisodicOracle(g["B"], g["N"], k=g["k"], use_c_scan=use_c)
    orc.memory[...] = g["mem0_full"]
    for t in range(g["steps"]):
        out = orc.step(g["q"][t], g["alpha"][t])
        assert np.array_equal(out["idx"], g["idx"][t])
        assert np.array_equal(out["d2"].view(np.uint32), g["d2"][t].view(np.uint32))
        assert _rel(out["reward"], g["reward"][t]).max() <= REL_TOL
        np.testing.assert_allclose(orc.ed_rms.mean, g["mean"][t], rtol=1e-6)
        np.testing.assert_allclose(orc.ed_rms.var, g["var"][t], rtol=1e-4, atol=1e-9)
        assert float(orc.ed_rms.count) == g["count"][t]
        # log-only trackers (episodic_novelty.py:82, intrinsic_novelty.py:32): the oracle's rewards differ from
        # the reference's by torch's non-correctly-rounded CPU sqrt (<= 3e-7), so do the statistics
        np.testing.assert_allclose(orc.epi_novel_rms.triple(), g["epi_rms"][t], rtol=2e-6)
        np.testing.assert_allclose(orc.intrinsic_novel_rms.triple(), g["intr_rms"][t], rtol=2e-6)
        assert orc.epi_novel_rms.triple()[2] == g["epi_rms"][t][2] and orc.intrinsic_novel_rms.triple()[2] == g["intr_rms"][t][2]
        orc.reset(g["done"][t])
    assert orc.cursor == int(g["cursor_final"])
    if g["mem_final"].size:
        assert np.array_equal(orc.memory, g["mem_final"])


@pytest.mark.parametrize("name", ["resets_mid_trajectory", "single_actor", "half_zero_slots"])
def test_torch_port_matches_reference_golden(name):
    import torch
    from oracle.torch_port import TorchCpuPort
    g = load_golden(name)
    port = TorchCpuPort(g["B"], g["N"], k=g["k"])
    port.memory.copy_(torch.from_numpy(g["mem0_full"]))
    for t in range(g["steps"]):
        out = port.step(torch.from_numpy(g["q"][t]), g["alpha"][t]).numpy()
        assert _rel(out, g["reward"][t]).max() <= 1e-6
        port.reset(g["done"][t])


def test_lane8_is_torch_cpu_order():
    """re-run the arithmetic probe on THIS host."""
    import torch
    from oracle.episodic_oracle import d2_lane8
    g = torch.Generator().manual_seed(5)
    mem = torch.randn(4000, 6, 32, generator=g)
    q = torch.randn(6, 32, generator=g)
    ref = ((mem - q.unsqueeze(0)) ** 2).sum(dim=-1).numpy()
    mine = d2_lane8(mem.numpy(), q.numpy()[None])
    assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


@pytest.mark.parametrize("n", [1, 5, 8, 64, 100, 129, 256, 300, 1000, 1024])
def test_pairwise_sum_is_numpy_order(n):
    from oracle.episodic_oracle import pairwise_sum_f32
    rng = np.random.default_rng(n)
    x = np.ascontiguousarray((rng.standard_normal((10, n)) * 3 + 5).astype(np.float32)).T   # (n,10), column-contiguous
    assert np.array_equal(pairwise_sum_f32(x).view(np.uint32), x.sum(axis=0).view(np.uint32))


def test_canonical_topk_tie_rule():
    from oracle.episodic_oracle import canonical_topk
    d2 = np.array([[1.0], [3.0], [3.0], [0.0], [3.0], [2.0]], np.float32)
    idx, v = canonical_topk(d2, 3, largest=True)
    assert idx[:, 0].tolist() == [1, 2, 4]
    idx, v = canonical_topk(d2, 3, largest=False)
    assert idx[:, 0].tolist() == [3, 0, 5]
    with pytest.raises(ValueError):
        canonical_topk(d2, 7)


def test_c_scan_equals_numpy_scan():
    from oracle import c_oracle
    from oracle.episodic_oracle import canonical_topk, d2_lane8
    rng = np.random.default_rng(3)
    for (N, B, k, largest) in [(777, 5, 10, True), (300, 7, 32, False), (10, 3, 10, True), (2000, 2, 1, True)]:
        mem = rng.standard_normal((N, B, 32), dtype=np.float32)
        mem[rng.random(N) < 0.3] = 0
        q = rng.standard_normal((B, 32), dtype=np.float32)
        i1, d1 = canonical_topk(d2_lane8(mem, q[None]), k, largest)
        for nt in (1, 3, 0):
            i2, d2 = c_oracle.scan_topk(mem, q, k, largest, nt)
            assert np.array_equal(i1, i2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


def test_multi_rank_mean_matches_single_rank_within_tolerance():
    """Block-partitioned actors (the reference's multi-rank MPI semantics, mpi_util.py:7-18)
    give the same rewards as one rank to far better than 1e-5."""
    from oracle.episodic_oracle import EpisodicOracle
    rng = np.random.default_rng(0)
    B, N = 16, 200
    a, b = EpisodicOracle(B, N), EpisodicOracle(B, N)
    for t in range(30):
        q = rng.standard_normal((B, 32), dtype=np.float32)
        r1 = a.step(q, None, n_ranks=1)["reward"]
        r4 = b.step(q, None, n_ranks=4)["reward"]
        assert _rel(r4, r1).max() < 1e-6


@pytest.mark.parametrize("B", [1, 6, 64, 256, 300])
def test_rnd_tail_oracle_matches_reference_golden(B):
    """lifelong_novelty.py:53-57 restated (oracle.LifelongTailOracle) vs the reference's own output."""
    import os
    from oracle.episodic_oracle import LifelongTailOracle
    from tests.conftest import ROOT
    from tests.golden.make_golden import rnd_features
    g = np.load(os.path.join(ROOT, "tests", "golden_rnd", "rnd_tail.npz"))
    pred, targ = rnd_features(B, 12)
    o = LifelongTailOracle()
    for t in range(12):
        a = o.modulator(pred[t], targ[t])
        np.testing.assert_allclose(a, g[f"alpha_{B}"][t], rtol=1e-12, atol=1e-12)
        assert abs(o.mean - g[f"mean_{B}"][t]) <= 1e-15 + 1e-12 * abs(g[f"mean_{B}"][t])
        assert abs(o.var - g[f"var_{B}"][t]) <= 1e-12 * abs(g[f"var_{B}"][t]) + 1e-18
        assert float(o.count) == g[f"count_{B}"][t]
