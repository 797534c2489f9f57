"""GPU tests of the device-resident sequence collector (include/ngu_seq.h, ngu_b200/sequence_buffer.py) through the
C ABI: against the golden recorded from the reference's own NGUAgent.collect_sequence / push_collected at
n_actors = 1, and against the oracle for many actors with episode ends, ring wrap of the store and ragged shapes.
Bars: byte moves and indices bit-exact; augmented / n-step rewards bit-exact (same float32 operations in the same
order, the discount powers come from the caller's framework)."""
import numpy as np
import pytest

from tests.test_sequence_cpu import NAMES, load_collect_golden, replay_golden

pytestmark = pytest.mark.gpu


def _hy(S, n, ov, cap):
    return dict(trace_length=S - 2, burnin_period=2, n_step=n, overlapping_period=ov, replay_capacity=cap)


def test_collector_matches_reference_golden(cuda_device):
    import torch
    from ngu_b200.sequence_buffer import SequenceCollector
    g = load_collect_golden()
    B, S, n, ov, _ = (int(v) for v in g["meta"])
    col = SequenceCollector(B, _hy(S, n, ov, 16), obs_shape=tuple(int(v) for v in g["obs_shape"]),
                            explr_beta=g["beta"], discounts=g["disc_pow"][:, 1:2], device=0)
    order = []

    def step(*a):
        col.step(*[torch.from_numpy(np.ascontiguousarray(x)) for x in a])
        order.extend(col.last_pushed().tolist())
    replay_golden(g, step)
    assert order == list(range(5)) and col.pushed == 5
    for nm in NAMES:
        got = col.store[nm][:5].cpu().numpy()
        want = g["push_" + nm]
        if got.dtype == np.float32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), nm
        else:
            assert np.array_equal(got, want), nm
    assert np.array_equal(col.local_mem_idx.cpu().numpy(), g["final_idx"])
    for nm in NAMES[:5]:
        assert np.array_equal(getattr(col, nm).cpu().numpy(), g["final_" + nm]), nm
    col.close()


@pytest.mark.parametrize("B,S,n,ov,obs_shape,cap,p_done", [(7, 12, 3, 5, (1, 5, 5), 9, 0.05), (64, 30, 5, 10, (1, 84, 84), 70, 0.02),
                                                           (300, 10, 2, 6, (3,), 301, 0.1), (2, 6, 1, 0, (1, 2, 2), 2, 0.0)])
def test_collector_vs_oracle(cuda_device, B, S, n, ov, obs_shape, cap, p_done):
    import torch
    from ngu_b200.sequence_buffer import SequenceCollector
    from oracle.sequence_oracle import SequenceOracle
    rng = np.random.default_rng(B + S)
    beta = rng.random((B, 1)).astype(np.float32)
    disc = (0.9 + 0.099 * rng.random((B, 1))).astype(np.float32)
    pw = torch.cat([torch.from_numpy(disc) ** i for i in range(n)], dim=1).numpy()
    col = SequenceCollector(B, _hy(S, n, ov, cap), obs_shape=obs_shape, explr_beta=beta, discounts=disc, device=0)
    orc = SequenceOracle(B, S, n, ov, obs_shape, beta, pw)
    positions = []
    steps = 4 * (S + n) + 7
    for t in range(steps):
        po = rng.standard_normal((B,) + obs_shape, dtype=np.float32)
        pa, ac = rng.integers(0, 18, (B, 1)), rng.integers(0, 18, (B, 1))
        ri, re = np.abs(rng.standard_normal((B, 1), dtype=np.float32)), rng.standard_normal((B, 1), dtype=np.float32)
        done = rng.random(B) < p_done
        orc.step(po, pa, ac, ri, re, done)
        # device tensors in: what the episodic engine hands over (r_int stays on the GPU)
        col.step(torch.from_numpy(po).cuda(), torch.from_numpy(pa).cuda(), torch.from_numpy(ac).cuda(),
                 torch.from_numpy(ri).cuda(), torch.from_numpy(re).cuda(), torch.from_numpy(done).cuda())
        positions.extend(col.last_pushed().tolist())
    assert col.pushed == len(orc.pushes) == len(positions) and col.pushed > cap // 2
    assert positions == [i % cap for i in range(len(positions))]          # ring order = push order
    store = {nm: col.store[nm].cpu().numpy() for nm in NAMES}
    actor = col.store["actor"].cpu().numpy()
    for i in range(max(0, len(positions) - cap), len(positions)):        # the sequences still in the ring
        p, pos = orc.pushes[i], positions[i]
        assert actor[pos] == p["actor"]
        for nm in NAMES:
            got = store[nm][pos]
            if got.dtype == np.float32:
                assert np.array_equal(got.view(np.uint32), p[nm].view(np.uint32)), (i, nm)
            else:
                assert np.array_equal(got, p[nm]), (i, nm)
    assert np.array_equal(col.local_mem_idx.cpu().numpy(), orc.local_mem_idx)
    for nm in NAMES[:5]:
        assert np.array_equal(getattr(col, nm).cpu().numpy(), getattr(orc, nm)), nm
    col.close()


def test_rewards_flow_from_the_engine_on_device(cuda_device):
    """The point of the row: r_int produced by the episodic engine goes into the sequence buffers without a host hop."""
    import torch
    from ngu_b200.engine import EpisodicEngine
    from ngu_b200.sequence_buffer import SequenceCollector
    B, N = 16, 500
    eng = EpisodicEngine(B, N, device=0)
    col = SequenceCollector(B, _hy(8, 2, 3, 64), obs_shape=(1, 4, 4), device=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    want = []
    for t in range(25):
        q = torch.randn(B, 32, device="cuda", generator=g)
        r_int, _ = eng.step(q, None)
        want.append(r_int.clone())
        col.step(torch.zeros(B, 1, 4, 4, device="cuda"), torch.zeros(B, 1, dtype=torch.int64, device="cuda"),
                 torch.zeros(B, 1, dtype=torch.int64, device="cuda"), r_int.view(B, 1), torch.zeros(B, 1, device="cuda"),
                 torch.zeros(B, dtype=torch.bool, device="cuda"))
    torch.cuda.synchronize()
    assert col.pushed == 3 * B            # L = 10: pushes at steps 10, 17 (overlap 3 -> continues at 4), 24
    first = col.store["intrinsic_reward"][:B, :, 0].cpu()
    assert torch.equal(first, torch.stack(want[:8], dim=1).cpu())
    eng.close()
    col.close()
