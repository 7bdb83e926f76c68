"""ctypes binding of libotn.so (the C ABI declared in include/otn.h).

There is NO CPU fallback: if the shared library is missing or cannot be loaded, every accelerated entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libotn.so")

MODEL_FULL, MODEL_LOCAL = 0, 1
METRIC_EUCLIDEAN, METRIC_CANONICAL = 0, 1
STATUS_NAN, STATUS_NOT_SPD, STATUS_NEG_DISCRIMINANT = 1, 2, 4
FLAG_DENSE, FLAG_BLOCKS, FLAG_FACTORED, FLAG_COMPLEX = 1, 2, 4, 8

_dp = C.c_void_p  # double* (host or device)
_ip = C.c_void_p  # int*


class TrParams(C.Structure):
    _fields_ = [("rho_trust", C.c_double), ("radius_init", C.c_double), ("maxradius", C.c_double), ("niter", C.c_int),
                ("tcg_maxiter", C.c_int), ("tcg_abstol", C.c_double), ("tcg_reltol", C.c_double)]


class TrReport(C.Structure):
    _fields_ = [("f_iter", _dp), ("rho", _dp), ("radius", _dp), ("n_cg", _ip), ("on_boundary", _ip), ("accepted", _ip),
                ("x_iter", _dp), ("radius_inout", _dp), ("status", _ip), ("hvp_total", C.c_longlong)]


# name -> (restype, argtypes); must list every symbol include/otn.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "otn_last_error": (C.c_char_p, []),
    "otn_version": (C.c_int, []),
    "otn_build_id": (C.c_char_p, []),
    "otn_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                             _dp, _dp, C.c_int]),
    "otn_create_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                _dp, _dp, C.c_int, C.c_int]),
    "otn_destroy": (None, [C.c_void_p]),
    "otn_set_target_index": (C.c_int, [C.c_void_p, _ip, C.c_int]),
    "otn_set_metric": (C.c_int, [C.c_void_p, C.c_int]),
    "otn_set_metric_general": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "otn_query": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                            C.POINTER(C.c_longlong)]),
    "otn_stream": (C.c_void_p, [C.c_void_p]),
    "otn_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "otn_launch_count": (C.c_longlong, []),
    "otn_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "otn_profile_read": (C.c_int, [C.c_void_p, _dp, C.c_void_p, _dp]),
    "otn_model": (C.c_int, [C.c_void_p, _dp, C.c_int, _dp]),
    "otn_cost": (C.c_int, [C.c_void_p, _dp, C.c_int, _dp]),
    "otn_cost_grad": (C.c_int, [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp]),
    "otn_triple": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int, _dp, _dp, _dp]),
    "otn_hvp_cached": (C.c_int, [C.c_void_p, _dp, C.c_int, _dp]),
    "otn_triple_submit": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_longlong)]),
    "otn_triple_wait": (C.c_int, [C.c_void_p, C.c_longlong]),
    "otn_ortho2super": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_kraus2superop": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "otn_release_scratch": (C.c_int, [C.c_int]),
    "otn_embed_local": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_compose": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_gemm_rows": (C.c_int, [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "otn_gemm_rows_bcast": (C.c_int, [_dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]),
    "otn_gemm_rows_ex": (C.c_int, [_dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   _dp, _dp, C.c_int, C.c_void_p]),
    "otn_gemm_rows_tiles": (C.c_int, [C.c_int, C.c_int]),
    "otn_peer_signal": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "otn_peer_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "otn_ipc_alloc": (C.c_int, [C.c_longlong, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]),
    "otn_ipc_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "otn_ipc_close": (C.c_int, [C.c_void_p, C.c_int]),
    "otn_ipc_free": (C.c_int, [C.c_void_p, C.c_int]),
    "otn_rowshard_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "otn_rowshard_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "otn_rowshard_detach": (C.c_int, [C.c_void_p]),
    "otn_rowshard_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "otn_expm": (C.c_int, [_dp, _dp, C.c_int, C.c_double, _dp, _dp, C.c_int]),
    "otn_super2ortho": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_random_normal": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, _dp, C.c_int]),
    "otn_random_stiefel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_kicked_starts": (C.c_int, [_dp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int]),
    "otn_expm_batch": (C.c_int, [_dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int]),
    "otn_liouvillian": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, C.c_int, _dp, _dp, C.c_int]),
    "otn_project": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_rgrad_map": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_int]),
    "otn_retract": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _ip, C.c_int]),
    "otn_canonical_dot": (C.c_int, [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "otn_retract_complex": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _ip, C.c_int]),
    "otn_connection": (C.c_int, [_dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_int]),
    "otn_frobenius": (C.c_int, [_dp, _dp, _dp, C.c_int, C.c_longlong, C.c_int, _dp, C.c_int]),
    "otn_tr_default_params": (None, [C.POINTER(TrParams)]),
    "otn_tr_run": (C.c_int, [C.c_void_p, _dp, C.c_int, C.POINTER(TrParams), C.POINTER(TrReport)]),
    "otn_gather": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, _dp]),
    "otn_comm_unique_id": (C.c_int, [C.c_char_p]),
    "otn_comm_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "otn_comm_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _dp, _dp]),
    "otn_comm_destroy": (None, [C.c_void_p]),
}

_lib = None


class OtnError(RuntimeError):
    pass


def load():
    """Load libotn.so (once).  Raises OtnError if it is missing -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OtnError(f"{LIB_PATH} not found: build it with `python -m opentn_b200.build` "
                       "(opentn_b200 has no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:  # e.g. libcudart missing
        raise OtnError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise OtnError(load().otn_last_error().decode())


def ptr(a):
    """Raw address of a numpy array (host) or of anything exposing `data_ptr()` (torch tensor, host or CUDA);
    None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(f"cannot take the address of {type(a)}")


def as_f64(a, shape=None):
    """C-contiguous float64 host array (a view when already so)."""
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        out = out.reshape(shape)
    return out
