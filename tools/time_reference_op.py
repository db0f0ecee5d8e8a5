#!/usr/bin/env python
"""Time the UNMODIFIED reference CUDA op (oracle/_ref, built for sm_100a) next to ours on the same B200:
the "beat this on the same box" bar of SURVEY.md 8(d).  Forward = chamfer_3D.forward (2 launches of
NmDistanceKernel), fwd+bwd adds chamfer_3D.backward; the reference's calc_dcd torch-op body is timed on top.
Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ured_b200 as ured  # noqa: E402
from oracle import build  # noqa: E402


def clouds(B, n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, n, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return (x / x.norm(dim=2).amax(1).view(B, 1, 1)).cuda()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ref_calc_dcd(ref, x, gt, alpha=1000, lam=1):
    """model_utils.py:13-58 over the reference op (forward only + the op's own backward for d loss/d dist)."""
    B, n_x, _ = x.shape
    n_gt = gt.shape[1]
    d1 = torch.zeros(B, n_gt, device="cuda"); d2 = torch.zeros(B, n_x, device="cuda")
    i1 = torch.zeros(B, n_gt, device="cuda", dtype=torch.int32); i2 = torch.zeros(B, n_x, device="cuda", dtype=torch.int32)
    ref.forward(gt, x, d1, d2, i1, i2)
    d1.requires_grad_(); d2.requires_grad_()
    e1, e2 = torch.exp(-d1 * alpha), torch.exp(-d2 * alpha)
    c1 = torch.zeros_like(i2); c1.scatter_add_(1, i1.long(), torch.ones_like(i1))
    w1 = (c1.gather(1, i1.long()).float() ** lam + 1e-6) ** (-1) * (n_gt / n_x)
    c2 = torch.zeros_like(i1); c2.scatter_add_(1, i2.long(), torch.ones_like(i2))
    w2 = (c2.gather(1, i2.long()).float() ** lam + 1e-6) ** (-1) * (n_x / n_gt)
    loss = ((1 - e1 * w1).mean(1) + (1 - e2 * w2).mean(1)) / 2
    g1, g2 = torch.autograd.grad(loss.sum(), [d1, d2])
    gx1, gx2 = torch.zeros_like(gt), torch.zeros_like(x)
    ref.backward(gt, x, gx1, gx2, g1.contiguous(), g2.contiguous(), i1, i2)
    return loss


def main():
    ref = build.load_ref()
    out = {"device": torch.cuda.get_device_name(0)}
    for name, B, n in [("cfg2", 640, 2048), ("cfg1", 32, 2048), ("cfg4", 16, 16384)]:
        x, gt = clouds(B, n, 1), clouds(B, n, 2)
        pairs = 2.0 * B * n * n
        d1 = torch.zeros(B, n, device="cuda"); d2 = torch.zeros(B, n, device="cuda")
        i1 = torch.zeros(B, n, device="cuda", dtype=torch.int32); i2 = torch.zeros(B, n, device="cuda", dtype=torch.int32)
        rec = {}
        if ref is not None:
            ms = timeit(lambda: ref.forward(gt, x, d1, d2, i1, i2), iters=5 if name != "cfg1" else 20)
            rec["reference_op_forward_ms"] = ms
            rec["reference_op_forward_tpair_s"] = pairs / ms / 1e9
            ms = timeit(lambda: ref_calc_dcd(ref, x, gt), iters=5 if name != "cfg1" else 20)
            rec["reference_calc_dcd_fwd_bwd_ms"] = ms
            rec["reference_calc_dcd_fwd_bwd_gpair_s"] = pairs / ms / 1e6
        ms = timeit(lambda: ured.nn_forward(gt, x), iters=20)
        rec["ours_forward_ms"] = ms
        rec["ours_forward_tpair_s"] = pairs / ms / 1e9

        def ours():
            xx, gg = x.detach().requires_grad_(), gt.detach().requires_grad_()
            ured.calc_dcd(xx, gg)[0].sum().backward()
        ms = timeit(ours, iters=20)
        rec["ours_calc_dcd_fwd_bwd_ms"] = ms
        rec["ours_calc_dcd_fwd_bwd_gpair_s"] = pairs / ms / 1e6
        out[name] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
