"""GPU tests of the drop-in classes through the reference's own call signatures
(IntrinsicNovelty.compute_intrinsic_novelty / reset_memory_if_done), driven exactly
like tests/golden/make_golden.py drives the reference: identity embedding, injected alpha."""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


class _NullLogger:
    def __init__(self):
        self.rows = []

    def log_scalar(self, name, value, step):
        self.rows.append((name, float(np.mean(value)), step))


@pytest.mark.parametrize("name", ["resets_mid_trajectory", "ring_wrap_n64", "half_zero_slots", "single_actor"])
def test_intrinsic_novelty_dropin_matches_reference_golden(cuda_device, name):
    import ngu_b200.pytorch_util as ptu
    from ngu_b200.engine import DEFAULT_HYPR
    from ngu_b200.intrinsic_novelty import IntrinsicNovelty
    g = load_golden(name)
    ptu.init_device()
    hy = dict(DEFAULT_HYPR, episodic_memory_capacity=g["N"], num_neighbors=g["k"])
    inn = IntrinsicNovelty(g["B"], 18, (1, 84, 84), hy, _NullLogger())
    inn.to(ptu.device)
    inn.epi_novel.embedding = lambda x: x                    # as make_golden.py does to the reference
    inn.epi_novel.episodic_memory = torch.from_numpy(g["mem0_full"])
    for t in range(g["steps"]):
        alpha = g["alpha"][t]
        inn.ll_novel.compute_lifelong_curiosity = lambda obs, a=alpha: torch.from_numpy(a.copy())
        out = inn.compute_intrinsic_novelty(torch.from_numpy(g["q"][t]))
        assert out.shape == (g["B"], 1) and out.device.type == "cpu"
        rel = np.abs(out.numpy()[:, 0].astype(np.float64) - g["reward"][t]) / np.abs(g["reward"][t])
        assert rel.max() <= 1e-5
        d2, idx = inn.epi_novel.knn()
        assert np.array_equal(idx.T, g["idx"][t])
        inn.reset_memory_if_done(torch.from_numpy(g["done"][t]))
    assert inn.epi_novel.memory_idx == int(g["cursor_final"])
    mem = inn.epi_novel.episodic_memory
    assert mem.shape == (g["N"], g["B"], 32)
    if g["mem_final"].size:
        assert np.array_equal(mem.cpu().numpy(), g["mem_final"])
    np.testing.assert_allclose(inn.epi_novel.ed_rms.mean, g["mean"][-1], rtol=1e-6)
    assert inn.epi_novel.ed_rms.count == g["count"][-1]
    # log-only trackers: running mean of the rewards (episodic_novelty.py:82, intrinsic_novelty.py:32)
    assert abs(inn.intrinsic_novel_rms.mean - g["reward"].mean()) <= 1e-4 * abs(g["reward"].mean())


def test_full_dropin_with_real_cnn_producers(cuda_device):
    """Observation-shaped input through Embedding + RND CNNs + engine; checks shapes, finiteness,
    the insert/reset bookkeeping, the logger hooks and the training step seam."""
    import ngu_b200.pytorch_util as ptu
    from ngu_b200.engine import DEFAULT_HYPR
    from ngu_b200.intrinsic_novelty import IntrinsicNovelty
    ptu.init_device()
    B, N = 6, 300
    log = _NullLogger()
    inn = IntrinsicNovelty(B, 18, (1, 84, 84), dict(DEFAULT_HYPR, episodic_memory_capacity=N), log)
    inn.to(ptu.device)
    g = torch.Generator().manual_seed(0)
    for t in range(5):
        obs = torch.randn(B, 1, 84, 84, generator=g).clamp_(-5, 5)
        r = inn.compute_intrinsic_novelty(obs)
        assert r.shape == (B, 1) and torch.isfinite(r).all() and (r >= 0).all()
        done = torch.zeros(B, dtype=torch.bool)
        done[t % B] = True
        inn.reset_memory_if_done(done)
        assert not inn.epi_novel.episodic_memory[:, t % B].any()
    assert inn.epi_novel.memory_idx == 5
    e = inn.epi_novel.compute_episodic_novelty(torch.randn(B, 1, 84, 84, generator=g))
    assert e.shape == (B,) and e.device.type == "cpu"
    obs = [torch.randn(B, 1, 84, 84, generator=g).to(ptu.device) for _ in range(2)]
    act = [torch.randint(0, 18, (B,), generator=g).to(ptu.device) for _ in range(2)]
    inn.step(obs, obs, act)
    names = {r[0] for r in log.rows}
    assert {"IntrinsicNoveltyMean", "EpisodicNoveltyMean", "KNeighborEucDistMean", "RNDLoss",
            "ActionpredictionLoss"} <= names
    sd = inn.epi_novel.state_dict()
    inn.epi_novel.load_state_dict(sd)
    assert inn.epi_novel.memory_idx == sd["engine"]["cursor"]
