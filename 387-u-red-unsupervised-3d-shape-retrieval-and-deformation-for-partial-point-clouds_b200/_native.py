"""ctypes binding of the C-ABI library (include/ured_chamfer.h).

This is the only place the package touches native code.  There is NO CPU fallback and no
alternative backend: if ``libured_chamfer.so`` is missing or a call fails, the caller gets an
exception (the reference silently ignores its op's return value, dist_chamfer_3D.py:45).
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libured_chamfer.so")
SRC_PATH = os.path.join(_HERE, "csrc", "ured_chamfer.cu")
INCLUDE_DIR = os.path.join(_ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

URED_FLAG_EXACT_ONLY = 1
URED_FLAG_NON_REG = 2
URED_FLAG_ONE_DIRECTION = 4
URED_FLAG_FP32_SCREEN = 8

_i, _u, _f, _p, _sz = ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_SIGNATURES = {
    # name: (restype, argtypes) -- device pointers travel as integers (c_void_p)
    "ured_abi_version": (_i, []),
    "ured_last_error_string": (ctypes.c_char_p, []),
    "ured_kernel_launches": (ctypes.c_ulonglong, []),
    "ured_packed_bytes": (_sz, [_i, _i]),
    "ured_pack_clouds": (_i, [_p, _i, _i, _p, _p, _p]),
    "ured_nn_scratch_bytes": (_sz, [_i, _i, _i]),
    "ured_nn_launch_shape": (_i, [_i, _i, _i, _u, _p, _p, _p, _p, _p, _p]),
    "ured_nn_packed": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _u, _p]),
    "ured_chamfer_workspace_bytes": (_sz, [_i, _i, _i]),
    "ured_chamfer_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _u, _p]),
    "ured_chamfer_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ured_nn_backward_one_direction": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "ured_dcd_forward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _f, _f, _f, _f, _u, _p, _p, _p, _p, _p, _p]),
    "ured_dcd_forward_ex": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _f, _f, _f, _f, _u, _p, _p, _p, _p, _p, _p, _f, _p]),
    "ured_dcd_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ured_merge_topk": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ured_topk_smallest": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "ured_emd_workspace_bytes": (_sz, [_i, _i]),
    "ured_emd_forward": (_i, [_p, _p, _i, _i, _f, _i, _p, _p, _p, _sz, _p]),
    "ured_emd_backward": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "ured_probe_ffma": (_i, [_p, _i, _i, ctypes.POINTER(ctypes.c_double), _p]),
    "ured_xchg_bytes": (_sz, [_i, _i, _i]),
    "ured_xchg_alloc": (_i, [_sz, ctypes.POINTER(ctypes.c_void_p)]),
    "ured_xchg_free": (_i, [_p]),
    "ured_xchg_export": (_i, [_p, _p]),
    "ured_xchg_import": (_i, [_p, ctypes.POINTER(ctypes.c_void_p)]),
    "ured_xchg_close": (_i, [_p]),
    "ured_xchg_status": (_i, [_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint), _p]),
    "ured_topk_exchange": (_i, [_p, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_void_p), _i, _i, _i, _p, _p, _u, _p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    """The CUDA extension is missing, cannot be loaded, or a call into it failed."""


def build_native(verbose=False):
    """Compile csrc/ured_chamfer.cu for sm_100a into libured_chamfer.so (in-tree)."""
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-o", LIB_PATH, SRC_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeLibraryError("nvcc failed:\n" + res.stdout + res.stderr)
    global _lib
    with _lock:
        _lib = None
    return res.stderr if verbose else LIB_PATH


def load():
    """Load the library once; raises NativeLibraryError if it is absent (never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:  # pragma: no cover - depends on the host
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.ured_abi_version() != 1:
            raise NativeLibraryError("ABI version mismatch between ured_chamfer.h and the built library")
        _lib = lib
    return _lib


def check(rc, what):
    """Turn a non-zero C-ABI return code into an exception."""
    if rc != 0:
        msg = load().ured_last_error_string().decode("utf-8", "replace")
        raise NativeLibraryError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
