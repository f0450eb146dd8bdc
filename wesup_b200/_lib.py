"""ctypes binding of libwesup_b200.so (the C ABI declared in include/wesup_b200.h).

There is NO fallback: if the shared library is missing, fails to load, or lacks
a symbol, importing the compute path raises.  torch is used only for device
memory (tensor.data_ptr()) and the current stream.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libwesup_b200.so"
ABI_VERSION = 7

F32, BF16 = 0, 1
CHW, HWC = 0, 1

_vp = c_void_p
_ip = POINTER(c_int)

# name -> (restype, argtypes); mirrors include/wesup_b200.h one to one
SIGNATURES = {
    "wesup_abi_version": (c_int, []),
    "wesup_last_error": (c_char_p, []),
    "wesup_kernel_launches": (ctypes.c_ulonglong, []),
    "wesup_hypercolumn_fwd": (c_int, [POINTER(_vp), _ip, _ip, _ip, c_int, c_int, c_int, _vp, c_int, c_int, _vp]),
    "wesup_hypercolumn_bwd_workspace_bytes": (c_size_t, [_ip, _ip, _ip, c_int, c_int, c_int]),
    "wesup_hypercolumn_bwd": (c_int, [_vp, c_int, c_int, _ip, _ip, _ip, c_int, c_int, c_int, POINTER(_vp), _vp, _vp]),
    "wesup_sp_stats_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "wesup_sp_stats": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wesup_sp_pool_fwd": (c_int, [_vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "wesup_sp_pool_bwd": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, _vp, c_int, c_int, _vp]),
    "wesup_hypercolumn_pool_fwd": (c_int, [POINTER(_vp), _ip, _ip, _ip, c_int, c_int, c_int, _vp, _vp, c_int, _vp, _vp]),
    "wesup_sp_pool_hypercolumn_bwd_workspace_bytes": (c_size_t, [_ip, _ip, _ip, c_int, c_int, c_int, c_int]),
    "wesup_sp_pool_hypercolumn_bwd": (c_int, [_vp, _vp, _vp, _ip, _ip, _ip, c_int, c_int, c_int, c_int, POINTER(_vp), _vp, _vp]),
    "wesup_levels_pool_fwd": (c_int, [POINTER(_vp), _ip, _ip, _ip, c_int, c_int, c_int, _vp, _vp, c_int, _vp, _vp]),
    "wesup_levels_pool_bwd_workspace_bytes": (c_size_t, [_ip, _ip, _ip, c_int, c_int, c_int]),
    "wesup_levels_pool_bwd": (c_int, [_vp, _vp, _vp, _ip, _ip, _ip, c_int, c_int, c_int, c_int, POINTER(_vp), _vp, _vp]),
    "wesup_footprint_bytes": (c_size_t, [_ip, _ip, c_int, c_int, c_int, c_int]),
    "wesup_footprint_build": (c_int, [_ip, _ip, c_int, c_int, c_int, c_int, _vp, _vp, _vp, _vp, c_int, _vp, _vp]),
    "wesup_levels_pool_fwd_fp": (c_int, [POINTER(_vp), _ip, _ip, _ip, c_int, c_int, c_int, _vp, _vp, c_int, _vp, _vp, _vp]),
    "wesup_levels_pool_bwd_fp": (c_int, [_vp, _vp, _vp, _ip, _ip, _ip, c_int, c_int, c_int, c_int, _vp, POINTER(_vp), _vp]),
    "wesup_hypercolumn_pool_fwd_walk": (c_int, [POINTER(_vp), _ip, _ip, _ip, c_int, c_int, c_int, _vp, _vp, c_int, _vp, _vp]),
    "wesup_sp_pool_hypercolumn_bwd_walk": (c_int, [_vp, _vp, _vp, _ip, _ip, _ip, c_int, c_int, c_int, c_int, POINTER(_vp), _vp, _vp]),
    "wesup_colsum_workspace_bytes": (c_size_t, [ctypes.c_long, c_int]),
    "wesup_colsum": (c_int, [_vp, ctypes.c_long, c_int, _vp, _vp, _vp]),
    "wesup_sp_paint": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp]),
    "wesup_label_propagate_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "wesup_label_propagate": (c_int, [_vp, c_int, c_int, c_int, _vp, c_int, c_float, _vp, _vp, _vp, _vp, _vp]),
    "wesup_label_propagate_exact": (c_int, [_vp, c_int, c_int, c_int, _vp, c_int, c_float, _vp, _vp, _vp, _vp, _vp]),
    "wesup_label_propagate_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "wesup_label_propagate_tc": (c_int, [_vp, c_int, c_int, c_int, _vp, c_int, c_float, _vp, _vp, _vp, _vp, _vp]),
    "wesup_label_propagate_tc_stats": (c_int, [_vp, c_int, c_int, POINTER(ctypes.c_ulonglong)]),
    "wesup_label_propagate_dev": (c_int, [_vp, c_int, c_int, _vp, _vp, c_int, c_float, _vp, _vp]),
    "wesup_upsample_sum": (c_int, [POINTER(_vp), _ip, _ip, c_int, c_int, c_int, c_int, c_int, _vp, c_int, _vp, _vp]),
    "wesup_slic_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "wesup_slic": (c_int, [_vp, c_int, c_int, c_int, c_int, c_double, c_int, c_int, _vp, _vp, _vp, _vp]),
    "wesup_slic_batch_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "wesup_slic_batch": (c_int, [_vp, c_int, c_int, c_int, c_int, c_int, c_double, c_int, c_int, _vp, _vp, _vp, _vp]),
    "wesup_slic_debug_times": (c_int, [_vp, c_int, c_int, c_int, c_int, POINTER(ctypes.c_ulonglong)]),
    "wesup_enforce_connectivity_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "wesup_enforce_connectivity": (c_int, [_vp, c_int, c_int, c_int, c_int, c_int, _vp, _vp, _vp, _vp]),
}

_lib = None


class WesupNativeError(RuntimeError):
    """Raised for any non-zero return of the C ABI (kept a RuntimeError so the
    reference trainer's per-iteration `except RuntimeError` still applies,
    /root/reference/models/base.py:234-237)."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m wesup_b200.build` "
                          "(there is no CPU or PyTorch fallback for the superpixel stage)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI/header mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.wesup_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libwesup_b200.so ABI {got} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wesup_last_error().decode(errors="replace")
        raise WesupNativeError(f"{what} failed (rc={rc}): {msg}")


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def ptr_array(ptrs):
    return (c_void_p * len(ptrs))(*[c_void_p(int(p)) for p in ptrs])
