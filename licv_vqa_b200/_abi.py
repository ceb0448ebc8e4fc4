"""ctypes binding of liblicv_b200.so (the C ABI declared in include/licv_b200.h).

There is no CPU path and no fallback: if the shared library is missing and cannot be built, or
the device is not a B200-class (sm_10x) GPU, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# LICV_LIB: an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("LICV_LIB") or os.path.join(_PKG, "lib", "liblicv_b200.so")

F32, BF16, F16 = 0, 1, 2
ROUND_Y, ROUND_NH, ROUND_NY, ROUND_T, ROUND_TEMPERED = 1, 2, 4, 8, 16
KD_KERNEL_GENERIC, KD_KERNEL_CLUSTER, KD_KERNEL_TMEM, KD_KERNEL_STREAM = 0, 1, 2, 3

_lib = None
_lock = threading.Lock()

_vp, _i64, _i32, _u32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint, C.c_float

# name -> (restype, argtypes); mirrors include/licv_b200.h declaration by declaration
SIGNATURES = {
    "licv_status_string": (C.c_char_p, [_i32]),
    "licv_abi_version": (_i32, []),
    "licv_device_info": (_i32, [C.POINTER(_i32)] * 3),
    "licv_icv_scale": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "licv_icv_scale_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "licv_inject_fwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _u32, _vp]),
    "licv_inject_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _u32, _vp]),
    "licv_inject_bwd_rows": (_i32, [_i64, _i32, _i32, _i32]),
    "licv_inject_bwd_spread": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _u32, _vp]),
    "licv_reduce_rows": (_i32, [_vp, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _vp]),
    "licv_icv_grad_finish": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _i32,
                                    _i32, _i32, _i32, _vp]),
    "licv_get_mask": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "licv_kd_prepare_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32,
                                    _vp, _vp, _vp, _vp]),
    "licv_kd_select_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32,
                                   _vp, _vp, _vp, _vp, _vp]),
    "licv_kd_loss_workspace_bytes": (_i64, [_i64]),
    "licv_kd_loss_plan": (_i32, [_i32, _i32, _f32, _i32, _i64]),
    "licv_debug_set_kd_stream": (None, [_i32]),
    "licv_kd_loss_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32,
                                    _i32, _f32, _vp, _vp, _i64, _i32, _i64, _i64, _i32, _u32, _vp]),
    "licv_kd_loss_fwd_bwd_dtemp": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32,
                                    _i32, _f32, _vp, _vp, _i64, _i32, _i64, _i64, _i32, _u32, _vp]),
    "licv_kd_loss_dtemp_workspace_bytes": (_i64, [_i64]),
    "licv_scale_inplace": (_i32, [_vp, _i64, _vp, _i32, _vp]),
    "licv_adamw_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32,
                               _i64, _f32, _f32, _vp, _vp, _vp]),
    "licv_adamw_step_partials": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32,
                                        _f32, _i64, _f32, _f32, _vp, _vp, _vp, _i32, _vp]),
    "licv_dp_region_bytes": (_i64, [_i64]),
    "licv_dp_region_alloc": (_i32, [_i64, C.POINTER(_vp), _vp]),
    "licv_dp_comm_create": (_i32, [C.POINTER(_vp), _i32, _i32, _vp, _vp, _i64]),
    "licv_dp_comm_destroy": (_i32, [_vp]),
    "licv_dp_comm_error": (_i32, [_vp]),
    "licv_dp_comm_reset_error": (_i32, [_vp]),
    "licv_dp_region_free": (_i32, [_vp]),
    "licv_dp_allreduce_adamw": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _f32, _f32,
                                       _f32, _f32, _f32, _i64, _f32, _vp, _vp, _vp]),
    "licv_host_session_create": (_i32, [C.POINTER(_vp), _i64, _i32]),
    "licv_host_session_destroy": (_i32, [_vp]),
    "licv_host_sync": (_i32, [_vp]),
    "licv_host_alloc_pinned": (_vp, [_i64]),
    "licv_host_free_pinned": (None, [_vp]),
    "licv_inject_fwd_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _u32]),
    "licv_inject_bwd_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _u32]),
    "licv_inject_fwd_host_save": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _u32]),
    "licv_inject_bwd_host_saved": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                          _u32]),
    "licv_kd_loss_fwd_bwd_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32,
                                         _f32, _i32, _f32, _vp, _i64, _i64, _i32, _i32, _u32]),
}


class LicvError(RuntimeError):
    """A liblicv_b200 call returned a non-zero status."""


def load(build_if_missing: bool = True):
    """Load (building first if needed) the shared library and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise ImportError(f"{LIB_PATH} is missing (run `python -m licv_vqa_b200.build`)")
            from . import build as _build
            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library out of sync
            fn.restype = res
            fn.argtypes = args
        if lib.licv_abi_version() != 1:
            raise ImportError("liblicv_b200.so ABI version mismatch: rebuild the library")
        _lib = lib
    return _lib


def status_string(rc: int) -> str:
    return load().licv_status_string(int(rc)).decode()


def check(rc: int, what: str = "liblicv_b200"):
    if rc != 0:
        raise LicvError(f"{what} failed: {status_string(rc)} (status {rc})")


def device_info():
    sm, major, minor = _i32(), _i32(), _i32()
    rc = load().licv_device_info(C.byref(sm), C.byref(major), C.byref(minor))
    return rc, sm.value, major.value, minor.value
