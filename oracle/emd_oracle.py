"""ctypes wrapper of oracle/emd_oracle.c -- TEST INFRASTRUCTURE ONLY (the CPU checker of the auction EMD)."""
import ctypes

import numpy as np

from . import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build.build_emd_oracle())
        lib.emd_oracle_forward.restype = ctypes.c_int
        lib.emd_oracle_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        _lib = lib
    return _lib


def emd_forward(xyz1, xyz2, eps, iters):
    """(dist [B,n] float32, assignment [B,n] int32, tie_events) for float32 clouds [B,n,3] (emd_cuda.cu:226-277)."""
    a = np.ascontiguousarray(xyz1, dtype=np.float32)
    b = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, n, _ = a.shape
    assert b.shape == a.shape
    dist = np.zeros((B, n), np.float32)
    assignment = np.zeros((B, n), np.int32)
    ties = ctypes.c_int(0)
    rc = _load().emd_oracle_forward(a.ctypes.data, b.ctypes.data, B, n, float(eps), int(iters), dist.ctypes.data, assignment.ctypes.data,
                                    ctypes.byref(ties))
    assert rc == 0
    return dist, assignment, ties.value
