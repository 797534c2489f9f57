"""ctypes binding of include/ngu_epi.h (libngu_epi.so).

This is the only place Python touches the native library.  There is no Python
or CPU fallback: if the shared object is missing the import fails loudly, and
every entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

MAX_K = 32
DIM = 32
MODE_REFERENCE, MODE_PAPER = 0, 1
ABI_VERSION = 5


class NguEpiCfg(C.Structure):
    _fields_ = [
        ("n_actors", C.c_int32), ("capacity", C.c_int32), ("dim", C.c_int32), ("k", C.c_int32),
        ("mode", C.c_int32), ("device", C.c_int32),
        ("kernel_epsilon", C.c_float), ("cluster_distance", C.c_float), ("pseudo_counts", C.c_float),
        ("max_similarity", C.c_float), ("max_reward_scale", C.c_float), ("reserved0", C.c_float),
        ("prior_count", C.c_double),
        ("scan_ctas_per_sm", C.c_int32), ("scan_variant", C.c_int32),
    ]


class NguEpiState(C.Structure):
    _fields_ = [
        ("ed_mean", C.c_double * MAX_K), ("ed_var", C.c_double * MAX_K), ("ed_count", C.c_double),
        ("epi_mean", C.c_double), ("epi_var", C.c_double), ("epi_count", C.c_double),
        ("intr_mean", C.c_double), ("intr_var", C.c_double), ("intr_count", C.c_double),
        ("cursor", C.c_int64), ("steps", C.c_int64), ("exchange_timeouts", C.c_int64),
        ("ll_mean", C.c_double), ("ll_var", C.c_double), ("ll_count", C.c_double),
    ]


# name -> (restype, argtypes); exactly the symbols include/ngu_epi.h declares
_P = C.c_void_p
SYMBOLS = {
    "ngu_epi_abi_version": (C.c_int, []),
    "ngu_epi_create": (C.c_int, [C.POINTER(NguEpiCfg), _P, C.POINTER(_P)]),
    "ngu_epi_destroy": (None, [_P]),
    "ngu_epi_last_error": (C.c_char_p, [_P]),
    "ngu_epi_memory": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "ngu_epi_scan": (C.c_int, [_P, _P, _P]),
    "ngu_epi_rank_sums": (C.c_int, [_P, C.POINTER(_P)]),
    "ngu_epi_finish": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "ngu_epi_log_update": (C.c_int, [_P, _P, _P]),
    "ngu_epi_append": (C.c_int, [_P, _P, _P]),
    "ngu_epi_reset": (C.c_int, [_P, _P, _P]),
    "ngu_epi_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ngu_epi_step_pipelined": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ngu_epi_step_rnd": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "ngu_epi_step_host": (C.c_int, [_P, _P, _P, _P, _P]),
    "ngu_epi_reset_host": (C.c_int, [_P, _P]),
    "ngu_epi_topk_host": (C.c_int, [_P, _P, _P]),
    "ngu_epi_get_state": (C.c_int, [_P, C.POINTER(NguEpiState)]),
    "ngu_epi_set_state": (C.c_int, [_P, C.POINTER(NguEpiState)]),
    "ngu_epi_set_cursor": (C.c_int, [_P, C.c_int64]),
    "ngu_epi_load_memory_host": (C.c_int, [_P, _P]),
    "ngu_epi_dump_memory_host": (C.c_int, [_P, _P]),
    "ngu_epi_p2p_export": (C.c_int, [_P, _P, C.c_int32]),
    "ngu_epi_p2p_connect": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "ngu_epi_log_flush": (C.c_int, [_P, _P]),
    "ngu_epi_launch_count": (C.c_int64, [_P]),
    "ngu_epi_scan_geometry": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "ngu_epi_debug_chunk": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "ngu_epi_debug_inject_timeout": (C.c_int, [_P]),
    "ngu_epi_plan_host": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), _P, C.c_int32]),
    "ngu_epi_debug_timeline": (C.c_int, [_P, _P, C.c_int32]),
    "ngu_epi_scan_timing_begin": (C.c_int, [_P, C.c_int32]),
    "ngu_epi_scan_timing_end": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
}



class NguSeqCfg(C.Structure):        # struct ngu_seq_cfg, include/ngu_seq.h
    _fields_ = [("n_actors", C.c_int32), ("seq_len", C.c_int32), ("n_step", C.c_int32), ("overlap", C.c_int32),
                ("obs_elems", C.c_int32), ("store_capacity", C.c_int32), ("device", C.c_int32), ("reserved0", C.c_int32)]


class NguSeqViews(C.Structure):      # struct ngu_seq_views: device pointers
    _fields_ = [(n, _P) for n in ("state", "prev_action", "curr_action", "intrinsic_reward", "extrinsic_reward",
                                  "local_mem_idx", "st_state", "st_prev_action", "st_curr_action", "st_intrinsic_reward",
                                  "st_extrinsic_reward", "st_augmented_reward", "st_nstep_reward", "st_actor", "counters",
                                  "last_pushed")]


# exactly the symbols include/ngu_seq.h declares
SEQ_SYMBOLS = {
    "ngu_seq_create": (C.c_int, [C.POINTER(NguSeqCfg), C.POINTER(_P)]),
    "ngu_seq_destroy": (None, [_P]),
    "ngu_seq_last_error": (C.c_char_p, [_P]),
    "ngu_seq_views_get": (C.c_int, [_P, C.POINTER(NguSeqViews)]),
    "ngu_seq_set_factors": (C.c_int, [_P, _P, _P]),
    "ngu_seq_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
}

# exactly the symbols include/ngu_ar.h declares
AR_HANDLE_BYTES = 128
AR_SYMBOLS = {
    "ngu_ar_create": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(_P)]),
    "ngu_ar_destroy": (None, [_P]),
    "ngu_ar_last_error": (C.c_char_p, [_P]),
    "ngu_ar_buffer": (C.c_int, [_P, C.POINTER(_P)]),
    "ngu_ar_export": (C.c_int, [_P, _P]),
    "ngu_ar_connect": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "ngu_ar_allreduce_avg": (C.c_int, [_P, _P]),
    "ngu_ar_timeouts": (C.c_int64, [_P]),
}

_lib = None


class NguEpiError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libngu_epi.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NguEpiError(
            f"{LIB_PATH} is missing. Build it with `python -m ngu_b200.build` "
            "(or __graft_entry__.build()). ngu_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SYMBOLS.items()) + list(SEQ_SYMBOLS.items()) + list(AR_SYMBOLS.items()):
        fn = getattr(lib, name)          # AttributeError if the ABI is out of date
        fn.restype = res
        fn.argtypes = args
    if lib.ngu_epi_abi_version() != ABI_VERSION:
        raise NguEpiError("libngu_epi.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().ngu_epi_last_error(handle)
        raise NguEpiError(f"libngu_epi error {rc}: {msg.decode() if msg else '?'}")
