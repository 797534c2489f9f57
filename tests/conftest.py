import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """Golden trajectory recorded from the unmodified reference (tests/golden/make_golden.py)."""
    from tests.golden.make_golden import prefill_memory
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    B, N, k, steps, seed = (int(v) for v in z["meta"])
    z.update(B=B, N=N, k=k, steps=steps, seed=seed)
    prefill = str(z["prefill"][0])
    mem0 = prefill_memory(prefill, N, B, seed, float(z["scale"]))
    if z["mem0"].size:
        assert mem0 is None and not z["mem0"].any() or np.array_equal(mem0, z["mem0"])
    z["mem0_full"] = mem0 if mem0 is not None else np.zeros((N, B, 32), np.float32)
    return z


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
