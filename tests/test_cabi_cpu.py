"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the
header declares, and fails loudly without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
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
    assert lib.ngu_epi_abi_version() == _cabi.ABI_VERSION == int(re.search(r"#define NGU_EPI_ABI_VERSION (\d+)", header).group(1))


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


@pytest.mark.parametrize("sched,taper", [(None, None), ("0.5,2,1", None), ("0.0,3,1", None), ("0.3,8,1,2", None), ("1.0,0", None),
                                         (None, "0.2,0.1"), ("0.5,8,1", "0.3,0.3"), ("0.0,4,1", "0.0,0.5"), (None, "0.9,0.0")])
def test_scan_schedule_partitions_the_tile_space_host_side(lib, monkeypatch, sched, taper):
    """Host logic of the scan kernel's work schedule (plan_geometry + chunk_range, no GPU needed): for
    many shapes, SM counts and kernel geometries every tile is handed out exactly once, the grid is one
    wave at most, and the statically owned chunks are the first ones."""
    import ctypes as C
    from ngu_b200._cabi import NguEpiCfg
    if sched:
        monkeypatch.setenv("NGU_EPI_SCHED", sched)
    else:
        monkeypatch.delenv("NGU_EPI_SCHED", raising=False)
    if taper:                         # rows of the dynamic matrix end in half / quarter chunks
        monkeypatch.setenv("NGU_EPI_TAPER", taper)
    else:
        monkeypatch.delenv("NGU_EPI_TAPER", raising=False)
    rng = np.random.default_rng(0)
    shapes = [(256, 30000), (1024, 30000), (1, 30000), (64, 5000), (37, 777), (3, 10), (5000, 33), (1, 10), (4096, 1000)]
    shapes += [(int(rng.integers(1, 600)), int(rng.integers(10, 40000))) for _ in range(12)]
    ctas = {0: 2, 1: 1, 2: 1, 3: 1, 4: 3, 5: 3, 6: 1, 7: 1}
    warps = {0: 8, 1: 8, 2: 8, 3: 16, 4: 8, 5: 4, 6: 8, 7: 4}
    for (B, N) in shapes:
        for sms, variant in [(148, 0), (132, 0), (148, 2), (148, 4), (20, 5)]:
            cfg = NguEpiCfg(n_actors=B, capacity=N, dim=32, k=10, mode=0, device=0, kernel_epsilon=1e-4,
                            cluster_distance=0.008, pseudo_counts=1e-3, max_similarity=8.0, max_reward_scale=5.0,
                            reserved0=0.0, prior_count=10.0, scan_ctas_per_sm=0, scan_variant=variant)
            geo = (C.c_int32 * 8)()
            assert lib.ngu_epi_plan_host(C.byref(cfg), sms, geo, None, 0) == 0
            grid, block, _, s1, _, tile_rows, c2, n_chunks = list(geo)
            chunks = (C.c_int32 * (2 * n_chunks))()
            assert lib.ngu_epi_plan_host(C.byref(cfg), sms, geo, chunks, n_chunks) == 0
            tpa = -(-N // tile_rows)
            cover = np.zeros(B * tpa + 1, np.int64)
            ch = np.frombuffer(chunks, dtype=np.int32).reshape(n_chunks, 2)
            assert (ch[:, 0] >= 0).all() and (ch[:, 0] <= ch[:, 1]).all() and (ch[:, 1] <= B * tpa).all()
            np.add.at(cover, ch[:, 0], 1)
            np.add.at(cover, ch[:, 1], -1)
            assert (np.cumsum(cover)[:-1] == 1).all(), (B, N, sms, variant, "tiles not covered exactly once")
            assert block == 32 * warps[variant] and 1 <= grid <= sms * ctas[variant]
            n_warps = grid * warps[variant]
            if c2 == 0:                                    # static schedule: one chunk per warp, nobody idle by design
                assert n_chunks <= n_warps and s1 >= 1
            else:                                          # dynamic: a full wave, the first n_warps chunks are the static ones
                assert grid == sms * ctas[variant] and c2 >= 1
                assert (ch[:min(n_warps, n_chunks), 1] - ch[:min(n_warps, n_chunks), 0] <= max(s1, c2)).all()
    assert lib.ngu_epi_plan_host(None, 148, None, None, 0) != 0
