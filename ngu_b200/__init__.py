"""ngu_b200 — B200-native episodic-novelty hot path for NGU (drop-in for minoring/NGU).

Only what the hot path needs lives here:
  csrc/                 sm_100a CUDA kernels + the C ABI (include/ngu_epi.h)
  _cabi.py, engine.py   ctypes binding and the per-GPU handle
  running_mean_std.py   host mirror of ngu/utils/mpi_util.py:RunningMeanStd
  episodic_novelty.py   drop-in EpisodicNovelty   (reference: episodic_novelty.py)
  intrinsic_novelty.py  drop-in IntrinsicNovelty  (reference: intrinsic_novelty.py)
  embedding.py, lifelong_novelty.py   the CNN producers (plain PyTorch/cuDNN)
  sharding.py           actor shards over the GPUs of one box + the rank-sum exchange
"""
__all__ = ["EpisodicEngine", "EpisodicNovelty", "IntrinsicNovelty", "RunningMeanStd"]


def __getattr__(name):
    if name == "EpisodicEngine":
        from .engine import EpisodicEngine
        return EpisodicEngine
    if name == "EpisodicNovelty":
        from .episodic_novelty import EpisodicNovelty
        return EpisodicNovelty
    if name == "IntrinsicNovelty":
        from .intrinsic_novelty import IntrinsicNovelty
        return IntrinsicNovelty
    if name == "RunningMeanStd":
        from .running_mean_std import RunningMeanStd
        return RunningMeanStd
    raise AttributeError(name)
