#!/usr/bin/env python
"""Record outputs of the UNMODIFIED reference auction-EMD op (oracle/_ref/emd, built by oracle/build.py::build_ref_emd) on a
B200 -> tests/golden/emd_ref_cuda_b200.npz.  Run on the GPU box:  python tests/golden/make_golden_emd_gpu.py [out.npz]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402

ref = ref_cuda.load_emd()
assert ref is not None, "oracle/_ref/emd not built"
out = {}
for name, (B, n, eps, iters, seed) in {"n1024_50": (4, 1024, 0.005, 50, 0), "n1024_200": (2, 1024, 0.002, 200, 1),
                                       "n2048_50": (2, 2048, 0.005, 50, 2)}.items():
    g = torch.Generator().manual_seed(1000 + seed)
    a, b = torch.rand(B, n, 3, generator=g), torch.rand(B, n, 3, generator=g)
    dist, assignment = ref_cuda.emd_forward(ref, a.cuda(), b.cuda(), eps, iters)
    out[f"{name}/xyz1"], out[f"{name}/xyz2"] = a.numpy(), b.numpy()
    out[f"{name}/dist"], out[f"{name}/assignment"] = dist.cpu().numpy(), assignment.cpu().numpy()
    out[f"{name}/eps"], out[f"{name}/iters"] = np.float32(eps), np.int32(iters)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "emd_ref_cuda_b200.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})
