"""ctypes/numpy front end of oracle/chamfer_oracle.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from . import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build.build_oracle())
        fp = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        lib.oracle_chamfer_forward.argtypes = [ctypes.c_int] * 3 + [fp, fp, fp, fp, ip, ip]
        lib.oracle_chamfer_forward.restype = ctypes.c_int
        lib.oracle_chamfer_backward.argtypes = [ctypes.c_int] * 3 + [fp, fp, fp, fp, fp, fp, ip, ip]
        lib.oracle_chamfer_backward.restype = ctypes.c_int
        lib.oracle_chamfer_backward_f64.argtypes = [ctypes.c_int] * 3 + [fp, fp, dp, dp, fp, fp, ip, ip]
        lib.oracle_chamfer_backward_f64.restype = None
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def chamfer_forward(xyz1, xyz2):
    """(dist1 [B,N] f32, dist2 [B,M] f32, idx1 [B,N] i32, idx2 [B,M] i32) -- chamfer3D.cu:136-154."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = np.zeros((b, n), np.float32); dist2 = np.zeros((b, m), np.float32)
    idx1 = np.zeros((b, n), np.int32); idx2 = np.zeros((b, m), np.int32)
    _load().oracle_chamfer_forward(b, n, m, xyz1, xyz2, dist1, dist2, idx1, idx2)
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """(gradxyz1, gradxyz2) in float32, j-ascending accumulation -- chamfer3D.cu:155-195."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros_like(xyz1); g2 = np.zeros_like(xyz2)
    _load().oracle_chamfer_backward(b, n, m, xyz1, xyz2, g1, g2, _f32(graddist1), _f32(graddist2),
                                    np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32))
    return g1, g2


def chamfer_backward_f64(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """Same terms accumulated in float64: the tolerance reference for the GPU's atomics."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros(xyz1.shape, np.float64); g2 = np.zeros(xyz2.shape, np.float64)
    _load().oracle_chamfer_backward_f64(b, n, m, xyz1, xyz2, g1, g2, _f32(graddist1), _f32(graddist2),
                                        np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32))
    return g1, g2
