"""ctypes wrapper around oracle/raster_ref.c (test infrastructure; see that file's header: PARITY UNPINNED)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_FMA = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_raster.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "raster_ref.c")):
            build()
        _LIB = ctypes.CDLL(so)
        for name in ("oracle_rasterize_fim_wim", "oracle_rasterize_fim_wim_fast"):
            fn = getattr(_LIB, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def _lib_fma():
    global _LIB_FMA
    if _LIB_FMA is None:
        so = os.path.join(_HERE, "liboracle_raster_fma.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "raster_ref.c")):
            build()
        _LIB_FMA = ctypes.CDLL(so)
        for name in ("oracle_rasterize_fim_wim_fma", "oracle_rasterize_fim_wim_fast_fma"):
            fn = getattr(_LIB_FMA, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB_FMA


def rasterize_fim_wim(faces, image_size, near=0.1, far=100.0, fast=True, fma=False):
    """faces (bs, nf, 3, 3) float32 -> fim (bs, S, S) int32, wim (bs, S, S, 3) float32.

    Restates ``nr.rasterize_face_index_map_and_weight_map(faces, image_size, False)`` as called at
    /root/reference/iPERCore/tools/human_digitalizer/renders/nmr.py:337,356 (near/far are the upstream defaults).
    ``fast=False`` runs the literal per-pixel-over-all-faces definition; ``fma=True`` the nvcc -fmad=true contraction
    model of the same source (raster_ref.c header).
    """
    faces = np.ascontiguousarray(faces, dtype=np.float32)
    bs, nf = faces.shape[:2]
    assert faces.shape[2:] == (3, 3)
    fim = np.empty((bs, image_size, image_size), np.int32)
    wim = np.empty((bs, image_size, image_size, 3), np.float32)
    if fma:
        fn = _lib_fma().oracle_rasterize_fim_wim_fast_fma if fast else _lib_fma().oracle_rasterize_fim_wim_fma
    else:
        fn = _lib().oracle_rasterize_fim_wim_fast if fast else _lib().oracle_rasterize_fim_wim
    rc = fn(faces.ctypes.data, bs, nf, image_size, near, far, fim.ctypes.data, wim.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle rasteriser allocation failed")
    return fim, wim
