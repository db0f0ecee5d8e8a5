"""CPU: the screening invariant of nn_kernel<SCREEN> checked on an arithmetic-exact emulation
(oracle/screen_model.c): an unambiguous query always has its reference argmin inside the winning 32-candidate chunk,
and screen + exact re-check + exact fallback reproduces the reference argmin everywhere."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_clouds


@pytest.fixture(scope="module")
def model(oracle):
    lib = ctypes.CDLL(oracle.build.build_screen_model())
    fp = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    lp = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    lib.screen_check.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, lp]
    lib.screen_check.restype = ctypes.c_long

    def run(q, c):
        q = np.ascontiguousarray(q, np.float32); c = np.ascontiguousarray(c, np.float32)
        stats = np.zeros(3, np.int64)
        bad = lib.screen_check(len(q), q, len(c), c, stats)
        return bad, stats
    return run


CASES = {
    "unit_ball_2048": lambda: (make_clouds(1, 1, 2048, "S")[0], make_clouds(2, 1, 2048, "S")[0] * 0.97),
    "uniform_cube": lambda: (make_clouds(3, 1, 2000, "U")[0], make_clouds(4, 1, 1000, "U")[0]),
    "ragged_tail": lambda: (make_clouds(5, 1, 777, "S")[0], make_clouds(6, 1, 1301, "S")[0]),
    "offset_1000": lambda: (make_clouds(7, 1, 600, "S")[0] + 1000.0, make_clouds(8, 1, 900, "S")[0] + 1000.0),
    "mixed_scale": lambda: (make_clouds(9, 1, 500, "S")[0] * 1e-3, make_clouds(10, 1, 700, "S")[0] * 50.0),
    "tiny_1e-20": lambda: (make_clouds(11, 1, 300, "S")[0] * 1e-20, make_clouds(12, 1, 300, "S")[0] * 1e-20),
    "near_duplicates": lambda: (make_clouds(13, 1, 800, "S")[0],
                                make_clouds(13, 1, 800, "S")[0] + 1e-7 * torch.randn(800, 3, generator=torch.Generator().manual_seed(1))),
    "dense_8192": lambda: (make_clouds(14, 1, 1024, "S")[0], make_clouds(15, 1, 8192, "S")[0]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_screening_invariant(model, name):
    q, c = CASES[name]()
    bad, (ambiguous, violations, mismatches) = model(q.numpy(), c.numpy())
    assert violations == 0, f"{name}: unambiguous query whose argmin is outside the winning chunk"
    assert mismatches == 0 and bad == 0
    if name in ("unit_ball_2048", "uniform_cube", "dense_8192"):
        assert ambiguous < 0.05 * len(q), f"{name}: {ambiguous} ambiguous queries -- the filter would be useless"


def test_screening_invariant_lattice_and_duplicates(model):
    g = torch.Generator().manual_seed(7)
    lattice_q = torch.randint(0, 4, (700, 3), generator=g).float()
    lattice_c = torch.randint(0, 4, (900, 3), generator=g).float()
    bad, stats = model(lattice_q.numpy(), lattice_c.numpy())
    assert bad == 0 and stats[0] > 0  # masses of exact ties: ambiguous, resolved by the exact fallback
    x = make_clouds(3, 1, 600, "S")[0]
    bad, _ = model(x.numpy(), torch.cat([x, x[:300]]).numpy())
    assert bad == 0
    bad, stats = model(x.numpy(), x.numpy())
    assert bad == 0
