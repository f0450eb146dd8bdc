"""ctypes wrapper around oracle/slic_ref.c -- TEST INFRASTRUCTURE ONLY.

Parity unpinned (scikit-image absent): see the header of slic_ref.c.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libslic_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "slic_ref.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_SO))
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.slic_ref.restype = ctypes.c_long
        lib.slic_ref.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                 i32p, i32p, ctypes.POINTER(ctypes.c_double)]
        lib.slic_ref_grid.restype = ctypes.c_long
        lib.slic_ref_grid.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        lib.slic_ref_connectivity.restype = ctypes.c_long
        lib.slic_ref_connectivity.argtypes = [i32p, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                              ctypes.c_long, i32p]
        lib.slic_ref_lab.restype = None
        lib.slic_ref_lab.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_long,
                                     ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        _lib = lib
    return _lib


def grid(H: int, W: int, n_segments: int):
    step, start = ctypes.c_int(), ctypes.c_int()
    k = _load().slic_ref_grid(H, W, n_segments, ctypes.byref(step), ctypes.byref(start))
    return int(k), step.value, start.value


def slic(img_hwc: np.ndarray, n_segments: int, compactness: float = 10.0, max_iter: int = 10,
         enforce_connectivity: bool = True, return_aux: bool = False):
    """Same call shape as skimage.segmentation.slic(image, n_segments, compactness)."""
    img = np.ascontiguousarray(img_hwc, dtype=np.float32)
    H, W, C = img.shape
    assert C == 3
    k, _, _ = grid(H, W, n_segments)
    if k <= 0:
        raise ValueError("degenerate SLIC grid")
    labels = np.empty((H, W), np.int32)
    raw = np.empty((H, W), np.int32)
    cent = np.empty((k, 5), np.float64)
    i32p = ctypes.POINTER(ctypes.c_int32)
    n = _load().slic_ref(img.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), H, W, n_segments,
                         float(compactness), max_iter, int(enforce_connectivity),
                         labels.ctypes.data_as(i32p), raw.ctypes.data_as(i32p),
                         cent.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if n < 0:
        raise ValueError("slic_ref rejected its input")
    if return_aux:
        return labels.astype(np.int64), raw, cent, int(n)
    return labels.astype(np.int64)


def connectivity(seg: np.ndarray, min_size: int, max_size: int):
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    H, W = seg.shape
    out = np.empty_like(seg)
    i32p = ctypes.POINTER(ctypes.c_int32)
    n = _load().slic_ref_connectivity(seg.ctypes.data_as(i32p), H, W, min_size, max_size,
                                      out.ctypes.data_as(i32p))
    return out, int(n)


def lab(img_hwc: np.ndarray, compactness: float):
    img = np.ascontiguousarray(img_hwc, dtype=np.float32)
    out = np.empty(img.shape, np.float64)
    _load().slic_ref_lab(img.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                         img.shape[0] * img.shape[1], float(compactness),
                         out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out
