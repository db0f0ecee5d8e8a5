import ctypes, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import ured_b200 as ured
from conftest import make_clouds
lib = ured._native.load()
B, n = 640, 2048
x = make_clouds(1, B, n, "S").cuda(); y = (make_clouds(2, B, n, "S") * 0.97).cuda()
out = (ctypes.c_longlong * 16)()
for rep in range(3):
    ured.nn_forward(x, y)
    torch.cuda.synchronize()
    lib.ured_debug_tc_profile(out)
    v = list(out)
    items = max(v[0], 1)
    print(f"items {v[0]}  per item: build {v[1]/items:.0f}  wait_full {v[2]/items:.0f}  read+reduce {v[3]/items:.0f}  recheck+Abuild {v[4]/items:.0f}  total {v[5]/items:.0f} cycles")
