"""Device glue with the reference's names (ngu/utils/pytorch_util.py:4-26).

The reference keeps one module-level ``device`` that every model reads; the
drop-in keeps that convention so ``ptu.init_device()`` in a caller's main()
still selects the GPU.  Unlike the reference there is no CPU fallback: the
episodic path needs a B200.
"""
import numpy as np
import torch

device = None


def init_device(use_gpu=True, index=None):
    """pytorch_util.py:7-14.  ``use_gpu=False`` is accepted for signature parity but
    the episodic engine will refuse to run without CUDA."""
    global device
    if use_gpu and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device() if index is None else index)
    else:
        device = torch.device("cpu")
    return device


def current_device():
    if device is None:
        init_device()
    return device


def to_tensor(x):
    """pytorch_util.py:17-22: numpy / python scalar -> float32 tensor on ``device``."""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float().to(current_device())
    return torch.tensor(x).float().to(current_device())


def to_numpy(tensor):
    """pytorch_util.py:25-26."""
    return tensor.detach().cpu().numpy()
