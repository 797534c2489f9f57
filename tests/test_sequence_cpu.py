"""CPU tests of the sequence-collector oracle (oracle/sequence_oracle.py) against the golden recorded from the
reference's own NGUAgent.collect_sequence / push_collected (tests/golden_seq/collect_b1.npz), and of the C ABI
surface of include/ngu_seq.h (symbols only: no compute without a GPU)."""
import os
import re

import numpy as np

from tests.conftest import ROOT

NAMES = ("state", "prev_action", "curr_action", "intrinsic_reward", "extrinsic_reward", "augmented_reward", "nstep_reward")


def load_collect_golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden_seq", "collect_b1.npz")))


def replay_golden(g, step):
    """Feeds the recorded environment / reward stream through `step` the way collect_sequence does
    (ngu_agent.py:76-93: the buffers receive prev_obs / prev_act, which are zeroed where an episode ended)."""
    B = int(g["meta"][0])
    obs_shape = tuple(int(v) for v in g["obs_shape"])
    prev_obs = np.zeros((B,) + obs_shape, np.float32)
    prev_act = np.zeros((B, 1), np.int64)
    for t in range(int(g["meta"][4])):
        done = g["in_done"][t]
        step(prev_obs, prev_act, g["in_act"][t], g["in_intr"][t], g["in_extr"][t], done)
        prev_obs, prev_act = g["in_obs"][t].copy(), g["in_act"][t].copy()
        prev_obs[done] = 0.0                                   # _reset_prev_if_done, :343-347
        prev_act[done] = 0


def test_sequence_oracle_matches_reference_golden():
    from oracle.sequence_oracle import SequenceOracle
    g = load_collect_golden()
    B, S, n, ov, _ = (int(v) for v in g["meta"])
    orc = SequenceOracle(B, S, n, ov, tuple(int(v) for v in g["obs_shape"]), g["beta"], g["disc_pow"])
    replay_golden(g, orc.step)
    assert len(orc.pushes) == g["push_state"].shape[0] == 5
    for i, p in enumerate(orc.pushes):
        for nm in NAMES:
            want = g["push_" + nm][i]
            if nm in ("augmented_reward", "nstep_reward"):     # float32 sums in the reference's order: identical bits
                assert np.array_equal(p[nm].view(np.uint32), want.view(np.uint32)), (i, nm)
            else:
                assert np.array_equal(p[nm], want), (i, nm)
    assert np.array_equal(orc.local_mem_idx, g["final_idx"])
    for nm in NAMES[:5]:
        assert np.array_equal(getattr(orc, nm), g["final_" + nm]), nm


def test_ngu_seq_header_symbols_are_exported():
    from ngu_b200 import _cabi
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "ngu_seq.h")).read()
    declared = set(re.findall(r"\b(ngu_seq_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.SEQ_SYMBOLS), declared ^ set(_cabi.SEQ_SYMBOLS)
    for name in declared:
        getattr(lib, name)


def test_sequence_collector_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ngu_b200 import _cabi
    from ngu_b200.sequence_buffer import SequenceCollector
    with pytest.raises(_cabi.NguEpiError):
        SequenceCollector(1, dict(trace_length=8, burnin_period=4, n_step=2, overlapping_period=4))
