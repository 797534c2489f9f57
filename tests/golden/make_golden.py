#!/usr/bin/env python
"""Generate golden trajectories by executing the UNMODIFIED reference module.

Runs only in the build container (needs ``/root/reference``); the fixtures it
writes (``tests/golden/*.npz``) are committed and are what travels to the GPU
box.  The reference (minoring/NGU) has no tests or golden vectors of its own
(SURVEY.md section 4), so these files are the pin for both the oracle and the
CUDA path.

How the reference is driven:
  * ``mpi4py`` is absent here; a single-rank stand-in whose
    ``COMM_WORLD.Allreduce(send, recv, op)`` copies ``send`` into ``recv`` is
    put on ``sys.path`` first (exactly one-rank MPI semantics for
    ``ngu/utils/mpi_util.py:17``).
  * ``IntrinsicNovelty`` is constructed unmodified; its CNN producers are
    bypassed by ``epi_novel.embedding = identity`` (so ``obs`` is the (B,32)
    controllable state) and by injecting a known RND modulator ``alpha`` through
    ``ll_novel.compute_lifelong_curiosity``.
  * Recorded per step: reference rewards (f64 under numpy>=2), the values of the
    reference's own ``topk`` on d^2, canonical indices (stable descending sort of
    the reference's distance expression; torch's own tie order is unspecified),
    the ``ed_rms`` state and the two log-only trackers (``epi_novel_rms``,
    ``intrinsic_novel_rms``).

Usage:  python tests/golden/make_golden.py        (writes next to this file)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NGU_REFERENCE", "/root/reference")

SHIM = '''
class _Comm:
    def Allreduce(self, send, recv, op=None):
        recv[...] = send
class MPI:
    SUM = "sum"
    COMM_WORLD = _Comm()
'''


def _import_reference():
    shim_dir = tempfile.mkdtemp(prefix="mpi4py_shim_")
    os.makedirs(os.path.join(shim_dir, "mpi4py"))
    with open(os.path.join(shim_dir, "mpi4py", "__init__.py"), "w") as f:
        f.write(SHIM)
    sys.path[:0] = [shim_dir, REF]
    import torch
    import ngu.utils.pytorch_util as ptu
    ptu.init_device(use_gpu=False)
    from ngu.models.intrinsic_novelty import IntrinsicNovelty
    from ngu.models.model_hypr import model_hypr
    return torch, IntrinsicNovelty, model_hypr


class _NullLogger:
    def log_scalar(self, *a, **k):
        pass


# name, B, N, steps, seed, scale, p_done, prefill ("none"|"randn"|"half_zero"|"dups"), alpha ("ones"|"randn")
CASES = [
    ("first_steps_zero_memory", 4, 64, 12, 11, 1.0, 0.0, "none", "ones"),
    ("ring_wrap_n64",           3, 64, 200, 12, 1.0, 0.0, "none", "randn"),
    ("resets_mid_trajectory",   7, 50, 120, 13, 1.0, 0.05, "none", "randn"),
    ("single_actor",            1, 777, 40, 14, 1.0, 0.02, "randn", "randn"),
    ("n_equals_k",              5, 10, 30, 15, 1.0, 0.0, "randn", "ones"),
    ("half_zero_slots",         6, 500, 25, 16, 1.0, 0.0, "half_zero", "randn"),
    ("duplicates_across_rank_k", 4, 96, 25, 17, 1.0, 0.0, "dups", "ones"),
    ("scale_1e-3",              5, 300, 25, 18, 1e-3, 0.0, "randn", "randn"),
    ("scale_50",                5, 300, 25, 19, 50.0, 0.0, "randn", "randn"),
    ("b64_n5000",              64, 5000, 6, 20, 1.0, 0.0, "randn", "randn"),
    ("b256_n300",             256, 300, 8, 21, 1.0, 0.01, "randn", "randn"),
    # the headline shape itself (BASELINE.json configs[2]): two recorded steps, the 983 MB prefill is rebuilt from the seed
    ("b256_n30000",           256, 30000, 2, 22, 1.0, 0.0, "randn", "randn"),
]


def prefill_memory(prefill, N, B, seed, scale):
    """Initial episodic memory [N,B,32] f32 from numpy's PCG64 stream (platform
    independent, unlike torch's vectorised CPU normal_), so tests can rebuild
    the large prefills from (recipe, seed) instead of storing them."""
    if prefill == "none":
        return None
    rng = np.random.default_rng(seed + 1000)
    if prefill == "randn":
        return (rng.standard_normal((N, B, 32), dtype=np.float32) * np.float32(scale))
    if prefill == "half_zero":
        m = rng.standard_normal((N, B, 32), dtype=np.float32) * np.float32(scale)
        m[rng.random(N) < 0.5] = 0.0
        return m
    if prefill == "dups":
        # few distinct rows repeated many times: exact d^2 ties straddle rank k
        base = rng.standard_normal((6, B, 32), dtype=np.float32) * np.float32(scale)
        return base[rng.integers(0, 6, size=N)].copy()
    raise ValueError(prefill)


def run_case(torch, IntrinsicNovelty, model_hypr, name, B, N, steps, seed, scale, p_done, prefill, alpha_kind):
    hy = dict(model_hypr, episodic_memory_capacity=N)
    inn = IntrinsicNovelty(B, 18, (1, 84, 84), hy, _NullLogger())
    en = inn.epi_novel
    en.embedding = lambda x: x
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    k = hy["num_neighbors"]

    mem_init = prefill_memory(prefill, N, B, seed, scale)
    if mem_init is not None:
        en.episodic_memory = torch.from_numpy(mem_init.copy())
    mem0 = en.episodic_memory.numpy().copy()

    rec = {k_: [] for k_ in ("q", "alpha", "done", "reward", "d2", "idx", "mean", "var", "count", "epi_rms", "intr_rms")}
    for _ in range(steps):
        q = torch.randn(B, 32, generator=g) * scale
        alpha = np.ones(B) if alpha_kind == "ones" else 1.0 + 2.0 * rng.standard_normal(B)
        inn.ll_novel.compute_lifelong_curiosity = lambda obs, a=alpha: torch.from_numpy(a.copy())
        # the reference's own distance expression, episodic_novelty.py:51-52 (pre-sqrt)
        d2_all = ((en.episodic_memory - q.unsqueeze(0)) ** 2).sum(dim=-1)
        ref_topk = d2_all.topk(k, dim=0)                       # values only are well defined
        canon = torch.sort(d2_all, dim=0, descending=True, stable=True)
        assert torch.equal(canon.values[:k], ref_topk.values)
        out = inn.compute_intrinsic_novelty(q)                 # THE reference call
        done = rng.random(B) < p_done
        inn.reset_memory_if_done(torch.from_numpy(done))
        rec["q"].append(q.numpy().copy())
        rec["alpha"].append(alpha.copy())
        rec["done"].append(done.copy())
        rec["reward"].append(out.numpy()[:, 0].astype(np.float64))
        rec["d2"].append(ref_topk.values.numpy().copy())
        rec["idx"].append(canon.indices[:k].numpy().astype(np.int32))
        rec["mean"].append(np.broadcast_to(en.ed_rms.mean, (k,)).astype(np.float64).copy())
        rec["var"].append(np.broadcast_to(np.ravel(en.ed_rms.var), (k,)).astype(np.float64).copy())
        rec["count"].append(float(np.ravel(en.ed_rms.count)[0]))
        # the log-only trackers (episodic_novelty.py:82, intrinsic_novelty.py:32): (mean, var, count) after the step
        rec["epi_rms"].append(np.array([np.ravel(t)[0] for t in (en.epi_novel_rms.mean, en.epi_novel_rms.var,
                                                                 en.epi_novel_rms.count)], np.float64))
        rec["intr_rms"].append(np.array([np.ravel(t)[0] for t in (inn.intrinsic_novel_rms.mean, inn.intrinsic_novel_rms.var,
                                                                  inn.intrinsic_novel_rms.count)], np.float64))
    out = {k_: np.stack(v) for k_, v in rec.items()}
    out["mem0"] = mem0
    out["mem_final"] = en.episodic_memory.numpy().copy()
    out["cursor_final"] = np.int64(en.memory_idx)
    out["meta"] = np.array([B, N, k, steps, seed], dtype=np.int64)
    out["scale"] = np.float64(scale)
    out["prefill"] = np.array([prefill])
    out["versions"] = np.array([f"torch {torch.__version__}", f"numpy {np.__version__}"])
    # big prefilled memories are regenerated from the seed by the tests, not stored
    if mem0.nbytes > (1 << 20):
        out["mem0"] = np.zeros((0,), np.float32)
        out["mem_final"] = np.zeros((0,), np.float32)
        out["mem_final_sum"] = np.float64(en.episodic_memory.double().sum().item())
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: B={B} N={N} steps={steps} reward[0,:3]={out['reward'][0, :3]}")


# ---- the real-CNN drop-in: Embedding + RND stacks in front of the scan --------------------------------
CNN_CASE = dict(B=8, N=256, steps=16, seed=31, mem_scale=2.0)


def fill_cnn_weights(module, seed):
    """Deterministic weights for a CNN stack from numpy's PCG64 (platform independent, and the same call fills the
    reference's modules here and the drop-in's modules in the GPU test: their state_dict keys and shapes agree).
    Weights ~ N(0, 2 / fan_in), biases ~ N(0, 0.05^2): activations stay O(1) through the four layers."""
    import torch
    rng = np.random.default_rng(seed)
    sd = module.state_dict()
    for name in sorted(sd):
        t = sd[name]
        if not torch.is_floating_point(t):
            continue
        if t.dim() >= 2:
            fan_in = int(np.prod(t.shape[1:]))
            w = rng.standard_normal(tuple(t.shape)).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_in))
        else:
            w = rng.standard_normal(tuple(t.shape)).astype(np.float32) * np.float32(0.05)
        t.copy_(torch.from_numpy(w))
    return module


def cnn_observations(B, steps, seed):
    """Synthetic observations [steps,B,1,84,84]: N(0,1) clipped to +-5, what VecNormalize hands the agent
    (envs/utils.py:37, atari_env_hypr.py:13)."""
    rng = np.random.default_rng(seed + 500)
    return np.clip(rng.standard_normal((steps, B, 1, 84, 84), dtype=np.float32), -5.0, 5.0)


def run_cnn_case(torch, IntrinsicNovelty, model_hypr):
    """Golden for the whole drop-in call on OBSERVATIONS: the unmodified reference IntrinsicNovelty with its real
    Embedding and RND networks (fixed weights), CPU float32.  Recorded per step: the reference's embeddings
    (embedding.py:48-49), the RND modulator (lifelong_novelty.py:49-57), the returned intrinsic reward
    (intrinsic_novelty.py:20-34), and the k+1 largest d^2 with their canonical slots (the (k+1)-th is recorded so
    that a test can tell which rank orders are separated by more than the embedding error)."""
    B, N, steps, seed = (CNN_CASE[x] for x in ("B", "N", "steps", "seed"))
    hy = dict(model_hypr, episodic_memory_capacity=N)
    k = hy["num_neighbors"]
    torch.manual_seed(seed)
    inn = IntrinsicNovelty(B, 18, (1, 84, 84), hy, _NullLogger())
    with torch.no_grad():
        fill_cnn_weights(inn.epi_novel.embedding, seed + 1)
        fill_cnn_weights(inn.ll_novel.predictor, seed + 2)
        fill_cnn_weights(inn.ll_novel.target, seed + 3)
    en = inn.epi_novel
    # memory prefilled at the embeddings' scale (|q| ~ 10): with an empty memory every neighbour is a zero slot
    en.episodic_memory = torch.from_numpy(prefill_memory("randn", N, B, seed, CNN_CASE["mem_scale"]).copy())
    obs = cnn_observations(B, steps, seed)
    rec = {k_: [] for k_ in ("emb", "alpha", "reward", "d2", "idx", "ll_rms", "ed_mean")}
    real_ll = inn.ll_novel.compute_lifelong_curiosity
    for t in range(steps):
        o = torch.from_numpy(obs[t])
        with torch.no_grad():
            q = en.embedding(o).detach()
        d2_all = ((en.episodic_memory - q.unsqueeze(0)) ** 2).sum(dim=-1)
        canon = torch.sort(d2_all, dim=0, descending=True, stable=True)
        seen = {}
        def spy(x, seen=seen):
            seen["alpha"] = real_ll(x)
            return seen["alpha"].clone()
        inn.ll_novel.compute_lifelong_curiosity = spy
        out = inn.compute_intrinsic_novelty(o)                 # THE reference call, real CNNs
        rec["emb"].append(q.numpy().copy())
        rec["alpha"].append(seen["alpha"].numpy().astype(np.float64).copy())
        rec["reward"].append(out.numpy()[:, 0].astype(np.float64))
        rec["d2"].append(canon.values[:k + 1].numpy().copy())
        rec["idx"].append(canon.indices[:k + 1].numpy().astype(np.int32))
        rec["ll_rms"].append(np.array([float(inn.ll_novel.ll_rms.mean), float(inn.ll_novel.ll_rms.var),
                                       float(inn.ll_novel.ll_rms.count)], np.float64))
        rec["ed_mean"].append(np.broadcast_to(en.ed_rms.mean, (k,)).astype(np.float64).copy())
    out = {k_: np.stack(v) for k_, v in rec.items()}
    out["mem_final"] = en.episodic_memory.numpy().copy()
    out["meta"] = np.array([B, N, k, steps, seed], dtype=np.int64)
    out["versions"] = np.array([f"torch {torch.__version__}", f"numpy {np.__version__}"])
    os.makedirs(os.path.join(HERE, "..", "golden_cnn"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "..", "golden_cnn", "dropin_cnn.npz"), **out)
    print(f"dropin_cnn: B={B} N={N} steps={steps} |emb| ~ {np.abs(out['emb']).mean():.3f} alpha[-1,:3]={out['alpha'][-1, :3]} "
          f"reward[-1,:3]={out['reward'][-1, :3]}")


def rnd_features(B, steps):
    """Synthetic RND (predictor, target) features [steps,B,128] from numpy's platform-independent PCG64."""
    rng = np.random.default_rng(3100 + B)
    pred = (rng.standard_normal((steps, B, 128)) * 0.3).astype(np.float32)
    targ = (rng.standard_normal((steps, B, 128)) * 0.3 + 0.1).astype(np.float32)
    return pred, targ


def run_rnd_case(torch, model_hypr):
    """Golden for the scalar tail of LifelongNovelty.compute_lifelong_curiosity
    (lifelong_novelty.py:49-57): the unmodified reference method with its CNN forward replaced by
    known (predictor, target) features."""
    from ngu.models.intrinsic_novelty.lifelong_novelty import LifelongNovelty
    out = {}
    for B in (1, 6, 64, 256, 300):
        ll = LifelongNovelty((1, 84, 84), model_hypr, _NullLogger())
        steps = 12
        pred, targ = rnd_features(B, steps)
        alphas, means, vars_, counts = [], [], [], []
        for t in range(steps):
            ll.forward = lambda obs, t=t: (torch.from_numpy(pred[t]), torch.from_numpy(targ[t]))
            a = ll.compute_lifelong_curiosity(torch.zeros(B, 1))      # THE reference call
            assert a.dtype == torch.float64
            alphas.append(a.numpy().copy())
            means.append(float(ll.ll_rms.mean)); vars_.append(float(ll.ll_rms.var)); counts.append(float(ll.ll_rms.count))
        out[f"alpha_{B}"] = np.stack(alphas)      # inputs are rebuilt from rnd_features(B, steps) by the tests
        out[f"mean_{B}"], out[f"var_{B}"], out[f"count_{B}"] = np.array(means), np.array(vars_), np.array(counts)
    out["batches"] = np.array([1, 6, 64, 256, 300])
    out["versions"] = np.array([f"torch {torch.__version__}", f"numpy {np.__version__}"])
    np.savez_compressed(os.path.join(HERE, "..", "golden_rnd", "rnd_tail.npz"), **out)
    print("rnd_tail: alpha[0,:3] =", out["alpha_6"][0, :3])


# ---- collect_sequence / push_collected buffering --------------------------------------------------------
def run_collect_case(torch, model_hypr):
    """Golden for the buffering half of NGUAgent.collect_sequence + push_collected (ngu_agent.py:56-94, :96-155): the
    UNMODIFIED reference methods executed on a stand-in `self` that carries exactly the attributes NGUAgent.__init__
    creates (:22-49; a small observation shape keeps the fixture small) and stub collaborators: an environment that
    returns seeded synthetic observations / rewards / episode ends, an intrinsic-novelty module that returns seeded
    rewards, TD errors of zero (the R2D2 networks are outside this path) and a replay memory that records what
    `push` receives.  n_actors = 1: the case in which the reference's advanced indexing is the per-actor write
    (include/ngu_seq.h).  R2D2Actor.compute_nstep_reward and NGUAgent.compute_priorities are the reference's own."""
    import types
    from ngu.models.ngu.ngu_agent import NGUAgent
    from ngu.models.r2d2.r2d2_actor import R2D2Actor
    B, obs_shape, calls, seed = 1, (1, 6, 6), 3, 41
    hy = dict(model_hypr)
    rng = np.random.default_rng(seed)
    rec = dict(obs=[], act=[], intr=[], extr=[], done=[])

    class Env:
        def reset(self):
            return torch.zeros((B,) + obs_shape)

        def step(self, action):
            o = torch.from_numpy(rng.standard_normal((B,) + obs_shape, dtype=np.float32))
            r = torch.from_numpy(rng.standard_normal((B, 1), dtype=np.float32))
            d = rng.random(B) < 0.01
            rec["act"].append(action.numpy().copy()); rec["obs"].append(o.numpy().copy())
            rec["extr"].append(r.numpy().copy()); rec["done"].append(d.copy())
            return o, r, d, None

    class Intr:
        def compute_intrinsic_novelty(self, obs):
            r = torch.from_numpy(np.abs(rng.standard_normal((B, 1), dtype=np.float32)))
            rec["intr"].append(r.numpy().copy())
            return r

        def reset_memory_if_done(self, done):
            pass

    class Memory:
        def __init__(self):
            self.pushed = []

        def push(self, *a):
            self.pushed.append([x.clone() if torch.is_tensor(x) else x for x in a])

    actor = types.SimpleNamespace(model_hypr=hy, explr_beta=torch.full((B, 1), 0.3), explr_beta_onehot=torch.zeros((B, 32)),
                                  discounts=torch.full((B, 1), 0.997), reset_hiddenstate_if_done=lambda d: None,
                                  policy=types.SimpleNamespace(get_hidden_state=lambda to_cpu=False: (None, None),
                                                               set_hidden_state=lambda h, c: None))
    actor.compute_nstep_reward = types.MethodType(R2D2Actor.compute_nstep_reward, actor)
    me = types.SimpleNamespace()
    me.envs, me.n_actors, me.n_act, me.obs_shape, me.model_hypr = Env(), B, 18, obs_shape, hy
    me.burnin_period = hy["burnin_period"]                                         # ngu_agent.py:22-27
    me.seq_len = hy["trace_length"] + me.burnin_period
    me.n_step = hy["n_step"]
    me.memory, me.r2d2_actor, me.intrinsic_novelty = Memory(), actor, Intr()
    me.init_hx, me.init_cx = torch.zeros((B, 4)), torch.zeros((B, 4))
    me.local_mem_idx = torch.zeros((B,), dtype=torch.int64)                        # :36-49
    me.mem_seq_len = me.seq_len + me.n_step
    me.mem_collected_mask = torch.full((B,), me.mem_seq_len - 1)
    me.state = torch.zeros((B, me.mem_seq_len) + obs_shape, dtype=torch.float32)
    me.prev_action = torch.zeros((B, me.mem_seq_len, 1), dtype=torch.int64)
    me.curr_action = torch.zeros((B, me.mem_seq_len, 1), dtype=torch.int64)
    me.intrinsic_reward = torch.zeros((B, me.mem_seq_len, 1), dtype=torch.float32)
    me.extrinsic_reward = torch.zeros((B, me.mem_seq_len, 1), dtype=torch.float32)
    me.prev_obs = me.envs.reset()
    me.prev_act = torch.zeros((B, 1), dtype=torch.int64)
    me.prev_ext_rew, me.prev_int_rew = torch.zeros((B, 1)), torch.zeros((B, 1))
    me.compute_td_error = lambda n, *a, **k: torch.zeros((me.seq_len - me.burnin_period, int(n), 1))
    for name in ("push_collected", "compute_priorities", "_reset_prev_if_done", "_reset_local_mem_if_done"):
        setattr(me, name, types.MethodType(getattr(NGUAgent, name), me))
    torch.manual_seed(seed)
    for _ in range(calls):
        NGUAgent.collect_sequence(me, random_policy=True)                          # THE reference call: 300 env steps each
    names = ("state", "prev_action", "curr_action", "intrinsic_reward", "extrinsic_reward", "augmented_reward", "nstep_reward")
    out = {"in_" + k: np.stack(v) for k, v in rec.items()}
    for j, nm in enumerate(names):
        out["push_" + nm] = np.stack([p[j].numpy()[0] for p in me.memory.pushed])
    out["final_idx"] = me.local_mem_idx.numpy().copy()
    for nm in names[:5]:
        out["final_" + nm] = getattr(me, nm).numpy().copy()
    out["beta"] = actor.explr_beta.numpy().copy()
    out["disc_pow"] = torch.cat([actor.discounts ** i for i in range(me.n_step)], dim=1).numpy().copy()
    out["meta"] = np.array([B, me.seq_len, me.n_step, hy["overlapping_period"], calls * 300], dtype=np.int64)
    out["obs_shape"] = np.array(obs_shape, dtype=np.int64)
    out["versions"] = np.array([f"torch {torch.__version__}", f"numpy {np.__version__}"])
    os.makedirs(os.path.join(HERE, "..", "golden_seq"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "..", "golden_seq", "collect_b1.npz"), **out)
    print(f"collect_b1: {calls * 300} steps, {len(me.memory.pushed)} sequences pushed, episode ends {int(out['in_done'].sum())}, "
          f"final idx {out['final_idx']}")


def lane8_probe(torch):
    """torch CPU's D=32 sum order must be the lane-8 order
    on the machine that produced the goldens."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.episodic_oracle import d2_lane8
    g = torch.Generator().manual_seed(99)
    mem = torch.randn(20000, 8, 32, generator=g)
    q = torch.randn(8, 32, generator=g)
    ref = ((mem - q.unsqueeze(0)) ** 2).sum(dim=-1).numpy()
    mine = d2_lane8(mem.numpy(), q.numpy()[None])
    frac = float((ref.view(np.uint32) == mine.view(np.uint32)).mean())
    print("lane-8 bit match vs torch CPU:", frac)
    assert frac == 1.0
    return frac


def main():
    torch, IntrinsicNovelty, model_hypr = _import_reference()
    lane8_probe(torch)
    only = sys.argv[1:]                       # optional: names of the cases to (re)generate
    for case in CASES:
        if not only or case[0] in only:
            run_case(torch, IntrinsicNovelty, model_hypr, *case)
    if not only or "rnd_tail" in only:
        os.makedirs(os.path.join(HERE, "..", "golden_rnd"), exist_ok=True)
        run_rnd_case(torch, model_hypr)
    if not only or "dropin_cnn" in only:
        run_cnn_case(torch, IntrinsicNovelty, model_hypr)
    if not only or "collect_b1" in only:
        run_collect_case(torch, model_hypr)


if __name__ == "__main__":
    main()
