"""ctypes binding of ``libheatflow_b200.so`` (the C-ABI declared in ``include/heatflow_b200.h``).

The product has no CPU fallback: if the shared library is missing, or no CUDA device is
usable, every entry point raises.  ``python __graft_entry__.py`` (or ``build()``) compiles it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libheatflow_b200.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/heatflow_b200.h line by line
SIGNATURES = {
    "hf_version": (C.c_int, []),
    "hf_last_error": (C.c_char_p, []),
    "hf_device_count": (C.c_int, []),
    "hf_create": (_vp, [C.c_int]),
    "hf_destroy": (None, [_vp]),
    "hf_set_ordering": (C.c_int, [_vp, _i32]),
    "hf_set_mesh": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hf_set_materials": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "hf_set_bcs": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "hf_set_bc_values": (C.c_int, [_vp, _vp]),
    "hf_build_operator": (C.c_int, [_vp, _f64, _i32]),
    "hf_get_sizes": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i64)]),
    "hf_get_csr": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "hf_set_state": (C.c_int, [_vp, _vp]),
    "hf_get_state": (C.c_int, [_vp, _vp]),
    "hf_set_source": (C.c_int, [_vp, _vp]),
    "hf_set_solver": (C.c_int, [_vp, _f64, _i32, _f64, _i32]),
    "hf_set_recycle": (C.c_int, [_vp, _i32]),
    "hf_get_solver_path": (C.c_int, [_vp]),
    "hf_set_sharing": (C.c_int, [_vp, _i32]),
    "hf_set_profile": (C.c_int, [_vp, _i32]),
    "hf_get_solve_profile": (C.c_int, [_vp, _vp, _vp]),
    "hf_step": (C.c_int, [_vp, _i32, _f64, _f64, _f64, C.POINTER(_i32), C.POINTER(_f64)]),
    "hf_get_rhs": (C.c_int, [_vp, _vp]),
    "hf_run": (C.c_int, [_vp, _i32, _vp, _f64, _f64, _i32, _vp, _vp, _vp, _vp]),
    "hf_sample": (C.c_int, [_vp, _i32, _vp, _vp]),
    "hf_get_stats": (C.c_int, [_vp, _vp]),
    "hf_debug_fx_shift": (C.c_int, [_vp, _i32]),
    "hf_debug_phase_times": (C.c_int, [_vp, _vp, _i32]),
    "hf_project_gradient": (C.c_int, [_vp, _vp, C.POINTER(_i32)]),
    "hf_spmv": (C.c_int, [_vp, _vp, _vp]),
    "hf_bench_kernels": (C.c_int, [_vp, _i32, _i32, _vp]),
    "hf_ens_create": (C.c_int, [_vp, _i32, _vp, _vp, _i32]),
    "hf_ens_run": (C.c_int, [_vp, _i32, _vp, _f64, _i32, _vp, _vp, _vp]),
    "hf_ens_get_state": (C.c_int, [_vp, _vp]),
    "hf_ens_get_path": (C.c_int, [_vp]),
    "hf_ens_destroy": (C.c_int, [_vp]),
}

_lib = None


class HeatflowError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"heatflow_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found - build it with `python __graft_entry__.py` "
            "(heatflow_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise HeatflowError(rc, load().hf_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Raw pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
