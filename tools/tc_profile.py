#!/usr/bin/env python
"""Phase counters and a per-query-tile timeline of nn_tc_kernel (needs a library built with -DURED_TC_PROFILE:
add the flag to NVCC_FLAGS in _native.py, rebuild, run this on the GPU, remove the flag).  cfg2-sized input."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import ured_b200 as ured  # noqa: E402
from conftest import make_clouds  # noqa: E402

lib = ured._native.load()
B, n = 640, 2048
x = make_clouds(1, B, n, "S").cuda()
y = (make_clouds(2, B, n, "S") * 0.97).cuda()
out = (ctypes.c_longlong * 16)()
trace = (ctypes.c_longlong * 96)()
for rep in range(3):
    ured.nn_forward(x, y)
    torch.cuda.synchronize()
    lib.ured_debug_tc_profile(out)
    v = list(out)
    items = max(v[0], 1)
    print(f"items {v[0]}  per item: build {v[1]/items:.0f}  wait_full {v[2]/items:.0f}  read+reduce {v[3]/items:.0f}  publish {v[4]/items:.0f}  "
          f"total {v[5]/items:.0f} | MMA warp 0: wait_empty {v[6]/items:.0f} fence {v[7]/items:.0f} issue+commit {v[8]/items:.0f} wait_operands {v[9]/items:.0f} cycles")
lib.ured_debug_tc_trace(trace)
t = [list(trace)[r * 24:(r + 1) * 24] for r in range(4)]
t0 = min(r[0] for r in t if r[0])
for name, r in zip(("reader set 0", "(fine)      ", "MMA warp 0 ", "resolvers   "), t):
    stamps = [v - t0 for v in r if v]
    if name.startswith("(fine)"):
        # reader set 0, fourth query tile of the item: per own tile [loop top, barrier seen, fence done, accumulator handed back]
        print("reader set 0, one query tile, per own tile: " + " | ".join(
            f"wait {stamps[i+1]-stamps[i]} fence {stamps[i+2]-stamps[i+1]} read {stamps[i+3]-stamps[i+2]}" + (f" gap {stamps[i+4]-stamps[i+3]}" if i + 4 < len(stamps) else "")
            for i in range(0, len(stamps) - 3, 4)))
        continue
    print(name, "start", stamps[0], "query-tile ends:", " ".join(str(b - a) for a, b in zip(stamps, stamps[1:])), "| last", stamps[-1])
