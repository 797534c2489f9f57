"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the
header declares, and fails loudly without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ngu_b200.build import build_library
    build_library()
    from ngu_b200 import _cabi
    return _cabi.load()


def test_exports_every_header_symbol(lib):
    from ngu_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "ngu_epi.h")).read()
    declared = set(re.findall(r"\b(ngu_epi_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ngu_epi_abi_version() == 3


def test_struct_layout_matches_header():
    from ngu_b200 import _cabi
    assert C.sizeof(_cabi.NguEpiCfg) == 64
    assert C.sizeof(_cabi.NguEpiState) == 8 * (32 + 32 + 1 + 6) + 24 + 24


def test_create_rejects_bad_config_and_missing_gpu(lib):
    import torch
    from ngu_b200 import _cabi
    h = C.c_void_p()
    bad = _cabi.NguEpiCfg(n_actors=4, capacity=64, dim=16, k=10)
    assert lib.ngu_epi_create(C.byref(bad), None, C.byref(h)) == -1
    assert b"dim" in lib.ngu_epi_last_error(None)
    bad = _cabi.NguEpiCfg(n_actors=4, capacity=5, dim=32, k=10)
    assert lib.ngu_epi_create(C.byref(bad), None, C.byref(h)) == -1       # k > capacity
    if not torch.cuda.is_available():
        ok = _cabi.NguEpiCfg(n_actors=4, capacity=64, dim=32, k=10)
        assert lib.ngu_epi_create(C.byref(ok), None, C.byref(h)) == -2    # NGU_ECUDA: no fallback
        assert b"no CPU fallback" in lib.ngu_epi_last_error(None)


def test_drop_in_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ngu_b200._cabi import NguEpiError
    from ngu_b200.engine import DEFAULT_HYPR
    from ngu_b200.intrinsic_novelty import IntrinsicNovelty

    class L:
        def log_scalar(self, *a):
            pass
    inn = IntrinsicNovelty(3, 18, (1, 84, 84), dict(DEFAULT_HYPR, episodic_memory_capacity=50), L())
    # attribute surface the reference's callers read
    assert inn.max_rew == 5.0 and inn.epi_novel.capacity == 50 and inn.epi_novel.controllable_state_dim == 32
    assert inn.epi_novel.memory_idx == 0 and inn.epi_novel.ed_rms.count == 10.0
    with pytest.raises(NguEpiError):
        inn.compute_intrinsic_novelty(torch.zeros(3, 1, 84, 84))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ngu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src, f


def test_running_mean_std_mirror_matches_oracle_restatement():
    import numpy as np
    from ngu_b200.running_mean_std import RunningMeanStd
    from oracle.episodic_oracle import RunningMeanStdOracle
    rng = np.random.default_rng(0)
    a, b = RunningMeanStd((10,)), RunningMeanStdOracle(10.0, shape=(10,))
    assert a.count == 10.0
    for _ in range(20):
        x = (rng.standard_normal((7, 10)) * 2 + 3).astype(np.float32)
        a.update(x)
        b.update(x)
    np.testing.assert_allclose(a.mean, b.mean, rtol=1e-6)
    np.testing.assert_allclose(a.var, b.var, rtol=1e-4)
    assert a.count == float(b.count)


def test_actor_partition():
    from ngu_b200.sharding import actor_partition
    assert actor_partition(256, 8) == [0, 32, 64, 96, 128, 160, 192, 224, 256]
    assert actor_partition(10, 4) == [0, 3, 6, 8, 10]
