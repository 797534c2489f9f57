"""ctypes wrapper of the plain-C oracle scan (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C

import numpy as np

from .build import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_scan_topk.restype = C.c_int
        lib.oracle_scan_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().oracle_max_threads())


def scan_topk(mem, q, k, largest=True, nthreads=0):
    """mem [N,B,32] f32 (reference layout), q [B,32] -> (idx int64 [k,B], d2 f32 [k,B]) best first."""
    mem = np.ascontiguousarray(mem, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    N, B, D = mem.shape
    assert D == 32 and q.shape == (B, 32)
    d2 = np.empty((k, B), np.float32)
    idx = np.empty((k, B), np.int32)
    rc = _load().oracle_scan_topk(mem.ctypes.data, q.ctypes.data, B, N, k, int(bool(largest)),
                                  d2.ctypes.data, idx.ctypes.data, int(nthreads))
    if rc != 0:
        raise ValueError("oracle_scan_topk: bad arguments (k > N?)")
    return idx.astype(np.int64), d2
