"""CPU model of nn_tc_kernel's operand construction (test infrastructure only; the product never imports oracle/).

The tensor-core screening pass feeds tcgen05.mma with bf16 operands that represent fp32 values EXACTLY: every fp32 value is
split into three bf16 pieces by repeated round-to-nearest-even (csrc/ured_chamfer.cu: bf16_split3), a candidate row holds
per coordinate [b1 b2 b1 b3 b2 b1 b3 b2] and [w1 w2 w3 0 ...], a query row [a1 a1 a2 a1 a2 a3 a2 a3] of -2q and [1 1 1 0 ...]
(tc_write_candidate / tc_write_query).  This module rebuilds those rows with numpy and evaluates the K = 32 product in float64,
so the CPU suite can check the algebra -- the pieces sum to the value, and the 27 products reproduce W_c - 2 q.c up to the
dropped a3.b3 terms -- independently of the GPU.  What the GPU adds on top is only the tensor core's fp32 accumulation
(measured by tools/microbench/tc_probe.cu, tests/test_gpu_tensor_screen.py).
"""
import numpy as np

PA = (0, 0, 1, 0, 1, 2, 1, 2)   # piece of -2q in slot t of a coordinate's eight products
PB = (0, 1, 0, 2, 1, 0, 2, 1)   # piece of c


def bf16_round(x):
    """float32 -> nearest bf16 (ties to even), returned as float32."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32)


def split3(x):
    """Three bf16 pieces (as float32) of every element; the residuals are computed in float32 like the kernel does."""
    x = np.asarray(x, np.float32)
    p1 = bf16_round(x)
    r1 = (x - p1).astype(np.float32)
    p2 = bf16_round(r1)
    r2 = (r1 - p2).astype(np.float32)
    p3 = bf16_round(r2)
    return p1, p2, p3


def query_rows(q):
    """[n, 3] float32 -> [n, 32] operand rows (float32 holding bf16 values)."""
    q = np.asarray(q, np.float32)
    rows = np.zeros((q.shape[0], 32), np.float32)
    for d in range(3):
        pieces = split3(np.float32(-2.0) * q[:, d])
        for t in range(8):
            rows[:, d * 8 + t] = pieces[PA[t]]
    rows[:, 24:27] = 1.0
    return rows


def candidate_rows(c):
    """[m, 3] float32 -> ([m, 32] operand rows, W [m]) with W = fma(z,z,fma(y,y,x*x)) in float32 like pack_kernel."""
    c = np.asarray(c, np.float32)
    x, y, z = (c[:, d].astype(np.float64) for d in range(3))
    w = np.float32(np.float32(x * x))                                    # mul.rn
    w = (y * y + w.astype(np.float64)).astype(np.float32)                # fma.rn: one rounding of the exact value
    w = (z * z + w.astype(np.float64)).astype(np.float32)
    rows = np.zeros((c.shape[0], 32), np.float32)
    for d in range(3):
        pieces = split3(c[:, d])
        for t in range(8):
            rows[:, d * 8 + t] = pieces[PB[t]]
    for t, piece in enumerate(split3(w)):
        rows[:, 24 + t] = piece
    return rows, w


def screen_scores(q, c):
    """The K = 32 product in float64 (every bf16 x bf16 product is exact in float64, the sum of 27 nearly so)."""
    a = query_rows(q).astype(np.float64)
    b, w = candidate_rows(c)
    return a @ b.astype(np.float64).T, w
