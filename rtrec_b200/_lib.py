"""ctypes binding of ``librtrec_b200.so`` (the C-ABI declared in ``include/rtrec_b200.h``).

There is no CPU fallback: importing the package works without a GPU (so the host-side logic can be
tested), but the first compute call raises :class:`RtrecB200Error` if the shared library or a CUDA
device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librtrec_b200.so")

RT_OK = 0
RT_ERR_ARG = -1
RT_ERR_CUDA = -2
RT_ERR_CAPACITY = -3
RT_ERR_NO_DEVICE = -4
RT_TOPK_DENSE = 0
RT_TOPK_SPARSE = 1


class RtrecB200Error(RuntimeError):
    pass


class FitConfig(C.Structure):
    _fields_ = [
        ("alpha", C.c_double),
        ("l1_ratio", C.c_double),
        ("tol", C.c_double),
        ("max_iter", C.c_int32),
        ("positive", C.c_int32),
        ("seed", C.c_uint32),
        ("nn", C.c_int32),
        ("n_samples", C.c_int32),
        ("nonneg", C.c_int32),
        ("rowmax_ptr", C.c_uint64),
        ("skip_trivial", C.c_int32),
    ]


_P = C.c_void_p
_I32, _I64, _U32, _F64 = C.c_int32, C.c_int64, C.c_uint32, C.c_double

# name -> (restype, argtypes); mirrors include/rtrec_b200.h one to one
PROTOTYPES = {
    "rt_version": (C.c_int, []),
    "rt_last_error": (C.c_char_p, []),
    "rt_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "rt_store_fold": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, _F64, _F64, _F64, _P, _P, _P, _I64, _F64, _I32, _I32,
                                _P, _P, _P, _I64, C.POINTER(_I64), C.POINTER(_F64), C.POINTER(_I32), C.POINTER(_I32), _P]),
    "rt_events_minmax": (C.c_int, [_P, _P, _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64),
                                   C.POINTER(_F64), _P]),
    "rt_events_item_stats": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P]),
    "rt_events_item_stats32": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P]),
    "rt_upload_events": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64),
                                   C.POINTER(_I64), C.POINTER(_F64), _I32, _P]),
    "rt_store_build": (C.c_int, [_P, _P, _P, _I64, _F64, _F64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P,
                                 C.POINTER(_I64), C.POINTER(C.c_int), _P]),
    "rt_store_lookup": (C.c_int, [_P, _P, _P, _I64, _P, _I64, _P, _P, _P, _P]),
    "rt_gram_rows": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _P, _P, _I64, _P]),
    "rt_gram": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _I64, _P]),
    "rt_gram_lower": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P, C.POINTER(_I32), _P]),
    "rt_gram_finish": (C.c_int, [_I32, _P, _I64, _P, _P, _P, _I64, _P]),
    "rt_gram_finish_rowmax": (C.c_int, [_I32, _P, _I64, _P, _P, _P, _I64, _P, C.POINTER(_I32), _P]),
    "rt_gram_finish_live": (C.c_int, [_I32, _P, _I64, _P, _P, _P, _I64, _P, C.POINTER(FitConfig), C.POINTER(_I32), C.POINTER(_I32), _P]),
    "rt_gram_block_rows": (C.c_int, [_I32, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "rt_gram_lower_blocks": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P]),
    "rt_gram_pull_cols": (C.c_int, [_I32, _P, _I32, _I32, _I64, _P]),
    "rt_gram_unpermute_rows": (C.c_int, [_I32, _I32, _P, _I64, _P, _P, _I64, _P]),
    "rt_gram_row_slots": (C.c_int, [_I32, _P, _I32, _P, _P]),
    "rt_gram_finish_p2p": (C.c_int, [_I32, _P, _I32, _I32, C.POINTER(_I32), _I64, _P, _P, _P, _I64, _I32, _P]),
    "rt_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P), _P]),
    "rt_ipc_open": (C.c_int, [_P, C.POINTER(_P)]),
    "rt_ipc_close": (C.c_int, [_P]),
    "rt_ipc_free": (C.c_int, [_P]),
    "rt_memset": (C.c_int, [_P, _I32, C.c_size_t, _P]),
    "rt_csr_split": (C.c_int, [_I32, _P, _P, _I32, _I32, _I32, _P, _P]),
    "rt_rng_table": (C.c_int, [_U32, _I64, _P, _P]),
    "rt_slim_solve_rows": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _I32, C.POINTER(FitConfig), _P, _P, _I64, _P, _P, _P, _P, _P,
                                     _I64, C.POINTER(_I64), _P, _P]),
    "rt_slim_solve": (C.c_int, [_P, _I64, _I32, _P, _I32, C.POINTER(FitConfig), _P, _P, _I64, _P, _P, _P, _P, _P, _I64,
                                C.POINTER(_I64), _P, _P]),
    "rt_slim_fit_pruned": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _I32, C.POINTER(FitConfig), _P, _I64, _P, _P,
                                     _P, _P, _I64, C.POINTER(_I64), _P, C.POINTER(_I32), C.POINTER(_I32), _P]),
    "rt_w_merge": (C.c_int, [_I32, _P, _P, _P, _I32, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _I64, C.POINTER(_I64), _P]),
    "rt_transpose": (C.c_int, [_I32, _I32, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "rt_slim_recommend": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "rt_score_tile": (C.c_int, [_I32, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "rt_w_pack_plan": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32),
                                 C.POINTER(_I64), _P]),
    "rt_w_pack_fill": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _I64, _P]),
    "rt_slim_recommend_packed": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32,
                                           _P, _P, _P, _P]),
    "rt_slim_recommend_candidates": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P]),
    "rt_topk_merge": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "rt_slim_similar": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P]),
    "rt_tc_pack_size": (C.c_int, [_I32, _I32, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64)]),
    "rt_tc_pack_build": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _I32, _P, _P, C.POINTER(_I32), _P]),
    "rt_values_bf16_exact": (C.c_int, [_P, _I64, C.POINTER(_I32), C.POINTER(_I32), _P]),
    "rt_slim_recommend_tc": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P,
                                       _P, _P, _P, _P, _P, _P, _P]),
    "rt_lru_replay": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _I64, _P, _P, C.POINTER(_I64)]),
    "rt_eval_metrics": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _P]),
    "rt_set_option": (C.c_int, [C.c_char_p, _I32]),
    "rt_gram_last_head": (_I32, []),
    "rt_release_scratch": (None, []),
    "rt_launch_count": (_I64, []),
    "rt_launch_count_reset": (None, []),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtrecB200Error(
            f"{LIB_PATH} is missing: build it with rtrec_b200/csrc/build.sh (or __graft_entry__.build()). "
            "rtrec_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != RT_OK:
        msg = load().rt_last_error()
        raise RtrecB200Error(f"{what or 'librtrec_b200'} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().rt_launch_count())
