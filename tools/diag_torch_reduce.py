#!/usr/bin/env python
"""Which summation order does torch's CUDA row reduction use?  Compares torch's x.sum(1) / x.mean(1) on the GPU, row by
row and bit for bit, with numpy float32 emulations of candidate schedules, and with the epilogue kernel's own sums.
Development tool behind the "Reduction order" contract of include/ured_chamfer.h."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ured_b200 as ured  # noqa: E402


def lanes_vec4(x, lanes=32, combine="seq", tree="up", vec=4):
    """thread t owns vectors t, t+lanes, ...; one accumulator per vector slot; slots combined; lanes combined by a tree."""
    B, n = x.shape
    per = lanes * vec
    full = n // per * per
    acc = np.zeros((B, lanes, vec), np.float32)
    xr = x[:, :full].reshape(B, -1, lanes, vec)
    for j in range(xr.shape[1]):
        acc = acc + xr[:, j]
    # remaining whole vectors, then the scalar tail to slot 0
    rem = x[:, full:]
    nv = rem.shape[1] // vec
    for t in range(nv):
        acc[:, t, :] = acc[:, t, :] + rem[:, t * vec:(t + 1) * vec]
    tail = rem[:, nv * vec:]
    for t in range(tail.shape[1]):
        acc[:, t, 0] = acc[:, t, 0] + tail[:, t]
    if combine == "seq":
        v = acc[:, :, 0]
        for i in range(1, vec):
            v = v + acc[:, :, i]
    else:
        v = (acc[:, :, 0] + acc[:, :, 1]) + (acc[:, :, 2] + acc[:, :, 3])
    v = v.astype(np.float32)
    L = lanes
    while L > 32:  # shared-memory tree down to one warp: thread t += thread t + L/2
        L //= 2
        v = (v[:, :L] + v[:, L:2 * L]).astype(np.float32)
    offs = [1, 2, 4, 8, 16] if tree == "up" else [16, 8, 4, 2, 1]
    for o in offs:
        w = v.copy()
        w[:, :32 - o] = v[:, :32 - o] + v[:, o:32]
        v = w.astype(np.float32)
    return v[:, 0]


def scalar_vt4(x, lanes=32):
    """non-vectorised schedule: thread t, unrolled 4 x stride."""
    B, n = x.shape
    acc = np.zeros((B, lanes, 4), np.float32)
    idx = 0
    while idx + lanes * 4 <= n:
        for i in range(4):
            acc[:, :, i] = acc[:, :, i] + x[:, idx + i * lanes: idx + (i + 1) * lanes]
        idx += lanes * 4
    i = 0
    while idx < n:
        m = min(lanes, n - idx)
        acc[:, :m, i] = acc[:, :m, i] + x[:, idx:idx + m]
        idx += m
        i += 1
    v = acc[:, :, 0]
    for i in range(1, 4):
        v = v + acc[:, :, i]
    for o in [1, 2, 4, 8, 16]:
        w = v.copy()
        w[:, :32 - o] = v[:, :32 - o] + v[:, o:32]
        v = w.astype(np.float32)
    return v[:, 0]


def main():
    torch.manual_seed(0)
    for B, n in [(16, 2048), (640, 2048), (37, 2000), (1000, 2048), (16, 1000), (64, 4096)]:
        x = (torch.rand(B, n) ** 4 * 0.01)
        xg = x.cuda()
        t_sum = xg.sum(1).cpu().numpy()
        t_mean = xg.mean(1).cpu().numpy()
        xn = x.numpy()
        zeros = torch.zeros(B, n, device="cuda")
        zi = torch.zeros(B, n, device="cuda", dtype=torch.int32)
        _, _, k_t = ured.retrieval.pair_scores(xg, zeros, zi, zi)          # cd_t = mean(x) + mean(0)
        k_t = k_t.cpu().numpy()
        print(f"B={B} n={n}: kernel cd_t vs torch mean: {(k_t != t_mean).sum()} rows differ; mean vs sum*f32(1/n): "
              f"{(t_mean != (t_sum * np.float32(1.0 / n))).sum()}")
        cands = {
            "32 lanes vec4 seq up (implemented)": lanes_vec4(xn),
            "32 lanes vec4 pairwise up": lanes_vec4(xn, combine="pair"),
            "32 lanes vec4 seq down": lanes_vec4(xn, tree="down"),
            "64 lanes vec4 seq up": lanes_vec4(xn, lanes=64),
            "128 lanes vec4 seq up": lanes_vec4(xn, lanes=128),
            "256 lanes vec4 seq up": lanes_vec4(xn, lanes=256),
            "512 lanes vec4 seq up": lanes_vec4(xn, lanes=512) if n >= 2048 else None,
            "32 lanes vec2": lanes_vec4(xn, vec=2, combine="seq"),
            "32 lanes vec8": lanes_vec4(xn, vec=8, combine="seq") if n % 256 == 0 else None,
            "32 lanes scalar vt4": scalar_vt4(xn),
        }
        for name, v in cands.items():
            if v is None:
                continue
            print(f"    {name:38s} vs torch sum: {(v != t_sum).sum():5d} rows differ   vs kernel: {((v * np.float32(1.0 / n)).astype(np.float32) != k_t).sum():5d}")
        bad = np.nonzero(lanes_vec4(xn) != t_sum)[0][:5]
        for r in bad:
            print(f"      row {r}: torch {t_sum[r]!r} emu {lanes_vec4(xn)[r]!r} f64 {xn[r].astype(np.float64).sum()!r}")


if __name__ == "__main__":
    main()
