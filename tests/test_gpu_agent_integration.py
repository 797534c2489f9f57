"""The drop-in inside the REAL caller: the reference's own ``NGUAgent.collect_sequence`` (ngu/models/ngu/ngu_agent.py:56-94,
unmodified, imported from the staged reference in oracle/_ref - test infrastructure) drives
``ngu_b200.intrinsic_novelty.IntrinsicNovelty`` through its real call sites: ``compute_intrinsic_novelty(next_obs)`` with
CPU observations (:74), the result written into the agent's float32 buffers (:79) and fed back as ``prev_int_rew`` (:89),
``reset_memory_if_done(done)`` (:86), ``push_collected`` with the reference's R2D2 networks (:81).  The only change to the
reference is the one INTEGRATION.md documents: the class the agent constructs (ngu_agent.py:8,33).

gym / ALE are not installed, so the vectorised environment is a seeded stand-in (observations ~ N(0,1) clipped to +-5 like
VecNormalize's output, random rewards and episode ends).  The same environment stream and the same random actions are
given to a second agent that keeps the reference's CPU ``IntrinsicNovelty`` with identical CNN weights; its return value is
cast to float32, which is what the reference returns under numpy 1.x - under numpy >= 2 the reference's own loop raises at
ngu_agent.py:79 (float64 into a float32 buffer), the drop-in's float32 return does not.

Bars: per-step intrinsic rewards within 1e-4 relative of the reference's (cuDNN float32 with TF32 off vs oneDNN float32
through three CNN stacks and the RND normalisation; the episodic part alone is checked to 1e-5 by the golden tests), the
agent's bookkeeping (local indices, sequences pushed to the replay memory) identical.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, N_SLOTS, SEED = 3, 400, 7


class _Logger:
    def log_scalar(self, *a, **k):
        pass


class _Env:
    """Seeded stand-in for the vectorised Atari environment (same stream for both agents)."""

    def __init__(self, n, seed):
        self.n, self.rng = n, np.random.default_rng(seed)

    def reset(self):
        return torch.zeros((self.n, 1, 84, 84))

    def step(self, action):
        o = torch.from_numpy(np.clip(self.rng.standard_normal((self.n, 1, 84, 84), dtype=np.float32), -5.0, 5.0))
        r = torch.from_numpy(self.rng.standard_normal((self.n, 1), dtype=np.float32))
        d = self.rng.random(self.n) < 0.004
        return o, r, d, None


def _make_agent(NGUAgent, hy, record):
    torch.manual_seed(SEED)                              # identical R2D2 / CNN initialisation on both sides
    agent = NGUAgent(_Env(B, SEED), B, 18, (1, 84, 84), hy, _Logger())
    inner = agent.intrinsic_novelty.compute_intrinsic_novelty

    def spy(obs):
        out = inner(obs).float()                         # (no-op for the drop-in; numpy-1 behaviour for the reference)
        record.append(out.numpy()[:, 0].copy())
        return out
    agent.intrinsic_novelty.compute_intrinsic_novelty = spy
    return agent


@pytest.mark.timeout(900)
def test_reference_ngu_agent_collects_with_the_dropin(cuda_device):
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref (the staged reference) is not present")
    _, RefIntrinsic, model_hypr, ref_ptu = ref_runner.import_reference(use_gpu=False)   # reference modules stay on the CPU
    import ngu.models.ngu.ngu_agent as agent_mod
    import ngu_b200.pytorch_util as ptu
    from ngu_b200.intrinsic_novelty import IntrinsicNovelty as DropIn
    ptu.init_device()
    hy = dict(model_hypr, replay_capacity=16, episodic_memory_capacity=N_SLOTS)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        want, got = [], []
        assert agent_mod.IntrinsicNovelty is RefIntrinsic
        ref_agent = _make_agent(agent_mod.NGUAgent, hy, want)
        agent_mod.IntrinsicNovelty = DropIn              # THE patch
        try:
            our_agent = _make_agent(agent_mod.NGUAgent, hy, got)
        finally:
            agent_mod.IntrinsicNovelty = RefIntrinsic
        assert isinstance(our_agent.intrinsic_novelty, DropIn)
        # same CNN weights on both sides (state-dict keys agree), drop-in on the GPU
        src, dst = ref_agent.intrinsic_novelty, our_agent.intrinsic_novelty
        dst.epi_novel.embedding.load_state_dict(src.epi_novel.embedding.state_dict())
        dst.ll_novel.predictor.load_state_dict(src.ll_novel.predictor.state_dict())
        dst.ll_novel.target.load_state_dict(src.ll_novel.target.state_dict())
        dst.to(ptu.device)
        for agent in (ref_agent, our_agent):
            torch.manual_seed(SEED + 1)                  # the same random actions (ngu_agent.py:62)
            agent.collect_sequence(random_policy=True)   # THE reference loop: 300 environment steps
        want, got = np.stack(want), np.stack(got)
        assert want.shape == got.shape == (300, B)
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert rel.max() <= 1e-4, f"intrinsic reward rel err {rel.max():.2e} at step {np.unravel_index(rel.argmax(), rel.shape)}"
        # the agent's own bookkeeping went the same way
        assert torch.equal(ref_agent.local_mem_idx, our_agent.local_mem_idx)
        assert ref_agent.memory.mem_pushed == our_agent.memory.mem_pushed >= 1
        n = min(ref_agent.memory.mem_pushed, hy["replay_capacity"])
        assert torch.allclose(ref_agent.memory.intrinsic_reward[:n], our_agent.memory.intrinsic_reward[:n], rtol=1e-4, atol=0)
        assert torch.allclose(ref_agent.memory.nstep_reward[:n], our_agent.memory.nstep_reward[:n], rtol=1e-3, atol=1e-3)
        assert torch.equal(ref_agent.memory.curr_action[:n], our_agent.memory.curr_action[:n])
        # and the episodic state the two modules hold agrees (the drop-in's view is the reference's (N, B, 32) layout)
        mem_ref = src.epi_novel.episodic_memory.numpy()
        mem_our = dst.epi_novel.episodic_memory.cpu().numpy()
        assert np.abs(mem_ref - mem_our).max() <= 1e-4
        assert src.epi_novel.memory_idx == dst.epi_novel.memory_idx == 300 % N_SLOTS
        dst.epi_novel._ensure_engine().close()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
