"""ctypes binding of `libglomar_b200.so` (C ABI declared in `include/glomar_b200.h`).

There is no CPU fallback: importing the compute classes on a machine without the built library, or
calling them without a CUDA device, raises.  Build the library with `python __graft_entry__.py`
(or `make -C glomargridding_b200/csrc`).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GLOMAR_B200_LIB selects another build of the same library (kernel experiments); never a fallback
LIB_PATH = os.environ.get("GLOMAR_B200_LIB") or os.path.join(_HERE, "libglomar_b200.so")

_lib = None


class GlomarB200Error(RuntimeError):
    """A libglomar_b200 entry point returned a non-zero status."""


def _declare(lib):
    p, i64, dbl, i32 = c_void_p, c_int64, c_double, c_int
    sigs = {
        "gmg_version": (c_int, []),
        "gmg_strerror": (c_char_p, [c_int]),
        "gmg_last_cuda_error": (c_int, []),
        "gmg_ellipse_ws_bytes": (c_int, [i64, POINTER(c_size_t)]),
        "gmg_ellipse_prepare_f64": (c_int, [p, p, p, p, p, p, i64, dbl, i32, p, c_size_t, p]),
        "gmg_ellipse_cov_f64": (c_int, [p, i64, i64, i64, dbl, dbl, i32, p, i64, i32, p]),
        "gmg_ellipse_pairs_f64": (c_int, [p, i64, p, p, i64, dbl, i32, p, p]),
        "gmg_ellipse_cov_p2p_f64": (c_int, [p, i64, dbl, dbl, i32, i32, i32, POINTER(c_int64), POINTER(c_void_p), i64,
                                            i32, p]),
        "gmg_ipc_alloc_f64": (c_int, [c_size_t, POINTER(c_void_p), p]),
        "gmg_ipc_free_f64": (c_int, [p]),
        "gmg_ipc_open_f64": (c_int, [p, POINTER(c_void_p)]),
        "gmg_ipc_close_f64": (c_int, [p]),
        "gmg_bessel_kv_f64": (c_int, [dbl, p, i64, i32, p, p]),
        "gmg_matern_f64": (c_int, [p, dbl, i32, p, i64, p, p]),
        "gmg_haversine_f64": (c_int, [p, p, i64, i64, i64, dbl, p, i64, p]),
    }
    u64 = c_uint64
    sigs.update({
        "gmg_gemm_f64": (c_int, [i32, i32, i64, i64, i64, dbl, p, i64, p, i64, dbl, p, i64, p]),
        "gmg_solve_ws_doubles": (c_int, [i64, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
        "gmg_potrf_batched_f64": (c_int, [p, i32, i64, p]),
        "gmg_potrs_batched_f64": (c_int, [p, i32, i64, i64, i32, p]),
        "gmg_potrf_step_f64": (c_int, [p, i32, i64, i32, i32, i32, i32, p]),
        "gmg_gesv_batched_f64": (c_int, [p, i32, i64, p]),
        "gmg_gesv_ws_doubles": (c_int, [i64, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
        "gmg_gesv_pivots_offset": (c_int64, [i64]),
        "gmg_gesv_blocked_f64": (c_int, [p, i32, i64, i64, p]),
        "gmg_krige_gather_batched_f64": (c_int, [p, i64, i64, p, i32, i64, p]),
        "gmg_krige_error_cov_lower_f64": (c_int, [p, i32, i64, p, i64, p]),
        "gmg_krige_solve_f64": (c_int, [p, i64, i64, p, p, i32, i64, p]),
        "gmg_krige_apply_batched_f64": (c_int, [p, i64, i64, p, i32, i64, i32, p]),
        "gmg_krige_weights_f64": (c_int, [p, i64, p, i64, p, i64, i64, p, i64, p]),
        "gmg_extended_inverse_f64": (c_int, [p, i64, i64, p, p, i64, p]),
        "gmg_normal_f64": (c_int, [p, i64, u64, u64, p]),
        "gmg_jacobi_eigh_ws_bytes": (c_int, [i64, POINTER(c_size_t)]),
        "gmg_jacobi_eigh_f64": (c_int, [p, i64, p, i64, i64, i32, dbl, p, c_size_t, p, p]),
        "gmg_cov2cor": (c_int, [p, i64, i64, i32, p, p, i64, p, p]),
        "gmg_cor2cov": (c_int, [p, i64, i64, i32, p, p, i64, p]),
        "gmg_uncompress": (c_int, [p, i64, p, i64, i32, dbl, dbl, p, i64, p]),
        "gmg_sym_eigvals_ws_bytes": (c_int, [i64, POINTER(c_size_t)]),
        "gmg_sym_eigvals_f64": (c_int, [p, i64, i64, p, p, c_size_t, p]),
        "gmg_variogram_f64": (c_int, [i32, p, i64, i64, i64, dbl, dbl, dbl, dbl, i32, i32, i32, dbl, p, i64, p]),
        "gmg_stationary_cov_f64": (c_int, [i32, p, p, i64, dbl, dbl, dbl, dbl, dbl, i32, i32, i32, dbl, p, i64, p]),
        "gmg_tril_f64": (c_int, [p, i64, i64, p]),
        "gmg_sigma_rot_f64": (c_int, [p, p, p, i64, p, p]),
        "gmg_det22_f64": (c_int, [p, i64, p, p]),
        "gmg_pair_haversine_f64": (c_int, [p, p, p, p, i64, dbl, p, p]),
        "gmg_pair_displacement_f64": (c_int, [p, p, p, p, i64, i32, dbl, p, p, p]),
        "gmg_nearest_index_f64": (c_int, [p, i64, p, i64, p, i64, i32, p]),
        "gmg_ellipse_fit_f64": (c_int, [p, i64, i64, p, p, p, i32, i32, i32, dbl, p, p, p, dbl, dbl, i32, i32, p, p, p, p,
                                        p, p, p]),
        "gmg_ellipse_nll_f64": (c_int, [p, i64, i64, p, p, p, i32, i32, i32, dbl, p, p, p, p]),
        "gmg_ellipse_train_f64": (c_int, [p, p, i64, p, i64, p, i64, i32, i32, i32, i32, dbl, dbl, i32, p, p, p, p, p,
                                          p]),
        "gmg_dist_unique_id": (c_int, [p]),
        "gmg_dist_init": (c_int, [p, i32, i32, POINTER(c_void_p)]),
        "gmg_dist_destroy": (c_int, [p]),
        "gmg_dist_ws_doubles": (c_int, [i64, i32, POINTER(c_int64)]),
        "gmg_potrf_solve_f64": (c_int, [p, p, i32, i64, i64, p, i64, p]),
        "gmg_host_pack_f64": (c_int, [p, i32, i64, i64, i64, p]),
        "gmg_host_pack_f32": (c_int, [p, i64, i64, i64, p]),
        "gmg_host_rows_symmetric": (c_int, [p, i32, i64, i64, i64, i64]),
        "gmg_microbench_fp64": (c_int, [i32, i32, i32, p, POINTER(c_double), POINTER(c_double), p]),
    })
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args


def lib():
    """The loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GlomarB200Error(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(glomargridding_b200 has no CPU fallback)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().gmg_strerror(status).decode()
        raise GlomarB200Error(f"{what} failed with status {status}: {msg}")


def exported_symbols() -> list[str]:
    """Every `gmg_*` symbol declared in include/glomar_b200.h (parsed from the header)."""
    import re

    header = os.path.join(os.path.dirname(_HERE), "include", "glomar_b200.h")
    with open(header) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(gmg_[a-z0-9_]+)\s*\(", text)))
