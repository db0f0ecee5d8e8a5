import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ured():
    """The package, with its C-ABI library built if this checkout has not been built yet."""
    import ured_b200 as pkg
    if not os.path.exists(pkg._native.LIB_PATH):
        pkg.build_native()
    pkg._native.load()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    from oracle import build, chamfer_oracle, torch_path
    build.build_oracle()

    class O:
        pass

    o = O()
    o.c = chamfer_oracle
    o.t = torch_path
    o.build = build
    return o


def make_clouds(seed, b, n, kind="U"):
    """Seeded CPU clouds: 'U' = torch.rand (the reference test's distribution), 'S' = unit-ball shape-like."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "U":
        return torch.rand(b, n, 3, generator=g)
    x = torch.randn(b, n, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    return x / x.norm(dim=2).amax(1).view(b, 1, 1)
