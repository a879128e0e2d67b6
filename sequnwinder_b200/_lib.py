"""ctypes loader for libsequnwinder_b200.so (the C ABI in include/sequnwinder_b200.h).

The library is built in-tree by ``sequnwinder_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no CPU fallback: if the shared object is missing this module raises, and every compute
entry point fails with ``SQW_ERR_NO_DEVICE`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SQW_LIB_PATH: development hook to A/B two builds of the same library
LIB_PATH = os.environ.get("SQW_LIB_PATH") or os.path.join(_HERE, "libsequnwinder_b200.so")

SQW_OK = 0
SQW_ERR_INVALID_ARG = -1
SQW_ERR_NO_DEVICE = -2
SQW_ERR_CUDA = -3
B200_SMS = 148   # 2 dies x 74 SMs (csrc/common.cuh kNumSMs)
SQW_ERR_BAD_BASE = -4
SQW_ERR_UNSUPPORTED = -5
SQW_ERR_NUMERIC = -6
SQW_ERR_NCCL = -7
SQW_ERR_STATE = -8

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)

# every symbol include/sequnwinder_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sqw_abi_version": (C.c_int, []),
    "sqw_last_error": (C.c_char_p, []),
    "sqw_device_count": (C.c_int, []),
    "sqw_set_device": (C.c_int, [C.c_int]),
    "sqw_get_device": (C.c_int, []),
    "sqw_num_kmers": (C.c_int64, [C.c_int, C.c_int]),
    "sqw_kmer_count": (C.c_int, [C.c_void_p, i64p, C.c_int64, C.c_int, C.c_int, i32p, u8p]),
    "sqw_kmer_count_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_admm_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "sqw_admm_destroy": (None, [C.c_void_p]),
    "sqw_admm_set_problem": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, f64p, i32p,
                                       f64p]),
    "sqw_admm_set_problem_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_admm_set_problem_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, u8p, f64p,
                                              f64p, i32p, f64p]),
    "sqw_admm_set_problem_counts_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                                  C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                  C.c_void_p]),
    "sqw_admm_set_structure": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, i32p, i32p, i32p,
                                         i32p, i32p, i32p, i32p]),
    "sqw_admm_set_params": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32]),
    "sqw_admm_set_ridge": (C.c_int, [C.c_void_p, C.c_double]),
    "sqw_admm_set_cta_budget": (C.c_int, [C.c_void_p, C.c_int32]),
    "sqw_nccl_unique_id": (C.c_int, [u8p]),
    "sqw_admm_set_comm": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, u8p]),
    "sqw_admm_execute": (C.c_int, [C.c_void_p, f64p, f64p]),
    "sqw_admm_get_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sqw_admm_get_log": (C.c_int, [C.c_void_p, C.c_int32, i32p, i32p, i32p, f64p, f64p, f64p,
                                   f64p, i32p, i32p]),
    "sqw_prepare_plan_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_double,
                                       C.c_void_p, i32p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_prepare_stats_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_prepare_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_double, C.c_int64, C.c_void_p, i32p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "sqw_prepare_apply_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_predict_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqw_hills_scan": (C.c_int, [C.c_void_p, i64p, C.c_int64, C.c_int, C.c_int, f64p, C.c_int, C.c_int,
                                 C.c_double, C.c_int, C.c_int32, i32p, i32p, f64p, i32p, u8p]),
    "sqw_hills_scan_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "sqw_admm_eval": (C.c_int, [C.c_void_p, f64p, f64p, C.c_double, f64p, f64p]),
    "sqw_lbfgs_rosenbrock": (C.c_int, [C.c_int32, f64p, C.c_double, i32p, i32p, i32p]),
}


class AdmmStats(C.Structure):
    _fields_ = [("outer_iters", C.c_int32), ("converged", C.c_int32), ("ran_admm", C.c_int32),
                ("total_admm_iters", C.c_int32), ("total_evals", C.c_int64),
                ("final_pho", C.c_double), ("final_fold", C.c_double),
                ("seconds_admm", C.c_double), ("seconds_total", C.c_double),
                ("launches", C.c_int64), ("seconds_eval", C.c_double),
                ("eval_bytes", C.c_int64), ("eval_mode", C.c_int32),
                ("ctas_per_block", C.c_int32), ("ring_stages", C.c_int32),
                ("tiles_per_cta", C.c_int32)]


class SqwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsequnwinder_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol. Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). sequnwinder_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != SQW_OK:
        msg = load().sqw_last_error()
        raise SqwError(rc, msg.decode(errors="replace") if msg else "")
