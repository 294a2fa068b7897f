"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/_build/libphilox_oracle.so."""

import ctypes as C
import os

import numpy as np

from .build_oracle import OUT, build

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT):
            build()
        _lib = C.CDLL(OUT)
        _lib.philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.b200mc_oracle_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    return _lib


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.empty(4, dtype=np.uint32)
    _load().philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def normals(seed: int, n_paths: int, n_steps: int, stream: int = 0, path_begin: int = 0) -> np.ndarray:
    """FP64 normals of the engine's documented stream for paths [path_begin, path_begin+n_paths)."""
    out = np.empty((n_paths, n_steps), dtype=np.float64)
    _load().b200mc_oracle_normals(seed & 0xFFFFFFFFFFFFFFFF, stream, path_begin, n_paths, n_steps, out.ctypes.data)
    return out
