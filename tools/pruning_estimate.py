#!/usr/bin/env python
"""CPU estimate (no GPU) of how much of the brute-force pair work a conservative spatial pruning could skip:
clouds Morton-sorted, candidates in 32-point chunks with an axis-aligned box, queries in warps of 32 Morton-adjacent
points with a box; a (warp, chunk) tile must be evaluated only if box-to-box distance^2 <= the warp's largest exact NN
distance^2 (an oracle bound: a real kernel only knows a running upper bound, so this is the best case)."""
import sys

import numpy as np
import torch

sys.path.insert(0, "tests")
from conftest import make_clouds  # noqa: E402


def morton(p, bits=10):
    q = ((p - p.min(0)) / (np.ptp(p, 0).max() + 1e-9) * (2 ** bits - 1)).astype(np.uint64)
    code = np.zeros(len(p), np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")


def estimate(nq, nc, kind, seed):
    q = make_clouds(seed, 1, nq, kind)[0].numpy().astype(np.float64)
    c = (make_clouds(seed + 1, 1, nc, kind)[0].numpy() * 0.97).astype(np.float64)
    q, c = q[morton(q)], c[morton(c)]
    d = ((q[:, None, :] - c[None, :, :]) ** 2).sum(-1)
    nn = d.min(1)
    needed = 0
    total = 0
    for w0 in range(0, nq, 32):
        qb_lo, qb_hi = q[w0:w0 + 32].min(0), q[w0:w0 + 32].max(0)
        bound = nn[w0:w0 + 32].max()
        for c0 in range(0, nc, 32):
            cb_lo, cb_hi = c[c0:c0 + 32].min(0), c[c0:c0 + 32].max(0)
            gap = np.maximum(0, np.maximum(cb_lo - qb_hi, qb_lo - cb_hi))
            total += 1
            needed += (gap ** 2).sum() <= bound
    return needed / total


if __name__ == "__main__":
    for nq, nc, kind in [(2048, 2048, "S"), (2048, 2048, "U"), (4096, 16384, "S")]:
        print(f"{kind} {nq}x{nc}: fraction of (warp, chunk) tiles that must be evaluated = {estimate(nq, nc, kind, 1):.3f}")
