"""Generate golden vectors from the UNMODIFIED reference CUDA op (oracle/_ref, built for sm_100a) on a B200.

Run on the GPU box (the reference sources are not needed, only the prebuilt module that oracle/build.py produced):
    gpurun -- 'python tests/golden/make_golden_gpu.py && cp tests/golden/chamfer_ref_cuda_b200.npz gpurun_out/'
Writes tests/golden/chamfer_ref_cuda_b200.npz: inputs (seeded, CPU-generated) and the op's dist1/dist2/idx1/idx2 and
gradients for the reference unit test's shapes and a few adversarial sets.  tests/test_oracle_pin.py then holds the C
oracle to these bits on the CPU, and tests/test_gpu_parity.py holds the B200 kernels to them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_clouds  # noqa: E402
from oracle import build  # noqa: E402


def main():
    ref = build.load_ref()
    assert ref is not None, "oracle/_ref is missing: run oracle/build.py where /root/reference exists"
    g = torch.Generator().manual_seed(7)
    cases = {
        "unit_test": (make_clouds(0, 4, 100, "U"), make_clouds(1, 4, 200, "U")),          # unit_test.py:15-16
        "timing": (make_clouds(10, 2, 2000, "U"), make_clouds(11, 2, 1000, "U")),         # unit_test.py:39-40
        "ragged_tail": (make_clouds(20, 2, 777, "S"), make_clouds(21, 2, 1301, "S")),
        "lattice": (torch.randint(0, 4, (2, 700, 3), generator=g).float(), torch.randint(0, 4, (2, 900, 3), generator=g).float()),
        "chair": (make_clouds(30, 1, 2048, "S"), make_clouds(31, 1, 2048, "S") * 0.97),
    }
    out = {}
    for name, (a, b) in cases.items():
        B, n, _ = a.shape
        m = b.shape[1]
        xa, xb = a.cuda(), b.cuda()
        d1 = torch.zeros(B, n, device="cuda"); d2 = torch.zeros(B, m, device="cuda")
        i1 = torch.zeros(B, n, device="cuda", dtype=torch.int32); i2 = torch.zeros(B, m, device="cuda", dtype=torch.int32)
        assert ref.forward(xa, xb, d1, d2, i1, i2) == 1
        gw = torch.Generator().manual_seed(2)
        w1, w2 = torch.randn(B, n, generator=gw), torch.randn(B, m, generator=gw)
        g1, g2 = torch.zeros_like(xa), torch.zeros_like(xb)
        assert ref.backward(xa, xb, g1, g2, w1.cuda(), w2.cuda(), i1, i2) == 1
        torch.cuda.synchronize()
        for k, v in dict(xyz1=a, xyz2=b, dist1=d1, dist2=d2, idx1=i1, idx2=i2, w1=w1, w2=w2, grad1=g1, grad2=g2).items():
            out[f"{name}_{k}"] = v.cpu().numpy()
    path = os.path.join(HERE, "chamfer_ref_cuda_b200.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes on", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()
