"""ctypes binding of include/lyssa_b200.h — the only door between the Python host code and the
CUDA kernels.  No fallback: importing this module without a built liblyssa_b200.so raises, and
every non-zero status becomes an exception carrying lys_last_error()."""
from __future__ import annotations

import ctypes
import os

from . import _build

c_int, c_i64, c_f, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

LYS_OK, LYS_EINVAL, LYS_ECUDA, LYS_EWORKSPACE, LYS_EUNSUPPORTED = 0, -1, -2, -3, -4
MAX_NONZERO, MAX_ATOMS, MAX_FEATURES = 32, 4096, 256
BOMP_SCREEN = 2
OMP_MAX_NONZERO = 64
COMM_HANDLE_BYTES = 64

# name -> (restype, argtypes); mirrors include/lyssa_b200.h declaration by declaration
SIGNATURES = {
    "lys_version": (c_int, []),
    "lys_last_error": (ctypes.c_char_p, []),
    "lys_build_fingerprint": (ctypes.c_char_p, []),
    "lys_device_info": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "lys_bomp_launch_count": (c_int, [c_int, c_int, c_i64, c_int]),
    "lys_profile_enable": (c_int, [c_int]),
    "lys_profile_fetch": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_char_p), c_int]),
    "lys_gram": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "lys_gram_workspace_bytes": (c_sz, [c_int, c_int]),
    "lys_gram_ws": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "lys_bomp_workspace_bytes": (c_sz, [c_int, c_int, c_i64, c_int]),
    "lys_bomp_encode": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_i64, c_int,
                                c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "lys_bomp_encode_ex": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_i64, c_int,
                                   c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_int, c_vp]),
    "lys_omp_workspace_bytes": (c_sz, [c_int, c_int, c_i64, c_int]),
    "lys_omp_encode": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_i64, c_int, c_f, c_int,
                               c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "lys_topk_select": (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "lys_thresh_workspace_bytes": (c_sz, [c_int, c_int, c_i64]),
    "lys_thresh_encode": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_int,
                                  c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "lys_iht_workspace_bytes": (c_sz, [c_int, c_int, c_i64]),
    "lys_iht_encode": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_int, ctypes.c_float, c_int,
                               c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "lys_corr_gemm": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_vp, c_int, c_vp]),
    "lys_bomp_encode_host": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_int,
                                     c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int]),
    "lys_codes_to_dense": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_i64, c_vp]),
    "lys_residual_workspace_bytes": (c_sz, [c_int, c_int, c_i64]),
    "lys_residual": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_i64, c_int,
                             c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_atom_csr_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "lys_build_atom_csr": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_ksvd_sweep_workspace_bytes": (c_sz, [c_int, c_int, c_i64, c_int]),
    "lys_approx_ksvd_sweep": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int,
                                      c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_ksvd_exact_workspace_bytes": (c_sz, [c_int, c_int]),
    "lys_ksvd_exact_sweep": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "lys_norm_cols": (c_int, [c_vp, c_i64, c_int, c_int, c_vp]),
    "lys_gather_cols": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_int, c_vp, c_i64, c_vp, c_vp]),
    "lys_odl_accumulate_workspace_bytes": (c_sz, [c_int, c_i64, c_int]),
    "lys_odl_accumulate": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_f, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_odl_update_workspace_bytes": (c_sz, [c_int, c_int]),
    "lys_odl_update_dict": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "lys_frobenius2": (c_int, [c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "lys_spm_total_cells": (c_int, [c_vp, c_int]),
    "lys_spm_pool": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, ctypes.c_float, c_vp, c_int, c_vp, c_int,
                             c_int, c_int, c_vp, c_vp, c_vp]),
    "lys_dsift_workspace_bytes": (c_sz, [c_int, c_int]),
    "lys_dsift_grid": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                               ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "lys_dsift": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp,
                          c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_dsift_batch": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lys_comm_create": (c_int, [c_int, c_int, ctypes.POINTER(c_vp)]),
    "lys_comm_export": (c_int, [c_vp, c_vp]),
    "lys_comm_connect": (c_int, [c_vp, c_vp]),
    "lys_comm_destroy": (c_int, [c_vp]),
}


class LyssaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("liblyssa_b200 status %d: %s" % (status, message))
        self.status = status


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing or stale) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LYSSA_B200_LIB")                              # override: bring-up builds only
    if not path:
        # build() compares the source fingerprint with the stamp of the in-tree .so and returns at once when they
        # agree, so a stale library (sources edited after the last build) is never loaded silently
        path = _build.build(force=bool(os.environ.get("LYSSA_B200_REBUILD")))
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().lys_last_error()
        raise LyssaError(status, msg.decode("utf-8", "replace") if msg else "")


def last_error() -> str:
    msg = load().lys_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
