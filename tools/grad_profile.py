#!/usr/bin/env python
"""Phase timestamps of grad_gather_kernel for one CTA (library built with -DURED_TC_PROFILE): cumulative cycles after the histogram,
the scans, the own terms and the gather.  cfg2- and cfg1-sized batches."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import ured_b200 as ured  # noqa: E402
from conftest import make_clouds  # noqa: E402

lib = ured._native.load()
out = (ctypes.c_longlong * 16)()
for B in (640, 32):
    x = make_clouds(1, B, 2048, "S").cuda().requires_grad_()
    y = (make_clouds(2, B, 2048, "S") * 0.97).cuda().requires_grad_()
    for rep in range(3):
        loss, _, _ = ured.calc_dcd(x, y)
        torch.cuda.synchronize()
        lib.ured_debug_tc_profile(out)
        loss.sum().backward()
        torch.cuda.synchronize()
        lib.ured_debug_tc_profile(out)
        v = list(out)
    print(f"B={B}: CTA 7 cumulative cycles: histogram {v[10]}  scans {v[11]}  own terms {v[12]}  gather {v[13]}")
