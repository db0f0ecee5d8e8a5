"""CPU: the auction-EMD oracle (oracle/emd_oracle.c) -- behaves like an auction, and is pinned to outputs of the reference's
own CUDA op recorded on a B200 (tests/golden/emd_ref_cuda_b200.npz, written by tests/golden/make_golden_emd_gpu.py)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emd_ref_cuda_b200.npz")


@pytest.fixture(scope="module")
def emd_oracle():
    from oracle import emd_oracle as eo
    return eo


def test_auction_converges_to_the_optimal_assignment(emd_oracle):
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    a, b = rng.random((2, 200, 3), dtype=np.float32), rng.random((2, 200, 3), dtype=np.float32)
    dist, assignment, _ = emd_oracle.emd_forward(a, b, 0.002, 3000)
    for s in range(2):
        cost = np.sqrt(((a[s][:, None, :] - b[s][None, :, :]) ** 2).sum(-1))
        r, c = linear_sum_assignment(cost)
        opt = cost[r, c].mean()
        got = np.sqrt(dist[s]).mean()
        assert len(set(assignment[s])) == 200                     # a bijection once the auction has run long enough
        assert opt - 1e-6 <= got <= opt + 0.002 + 1e-6            # eps-optimal: within eps per point of the optimum
    # few iterations: the last one assigns every remaining bidder to the object it bid on (no bijection), as in the reference
    dist, assignment, _ = emd_oracle.emd_forward(a, b, 0.005, 3)
    assert (assignment >= 0).all() and len(set(assignment[0])) < 200
    # dist is the squared distance to the assigned point
    s2 = ((a - np.take_along_axis(b, assignment[..., None].astype(np.int64), axis=1)) ** 2).sum(-1)
    assert np.allclose(dist, s2, rtol=1e-6)


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden vectors of the reference EMD op not recorded yet")
def test_oracle_reproduces_the_reference_op(emd_oracle):
    g = np.load(GOLDEN)
    checked = 0
    for case in sorted({k.split("/")[0] for k in g.files}):
        a, b = g[f"{case}/xyz1"], g[f"{case}/xyz2"]
        eps, iters = float(g[f"{case}/eps"]), int(g[f"{case}/iters"])
        for s in range(a.shape[0]):
            dist, assignment, ties = emd_oracle.emd_forward(a[s:s + 1], b[s:s + 1], eps, iters)
            if ties == 0:   # the reference's result is well defined (no bidders within its 1e-6 tolerance): bits must match
                assert np.array_equal(assignment[0], g[f"{case}/assignment"][s]), f"{case} pair {s}"
                assert np.array_equal(dist[0], g[f"{case}/dist"][s])
                checked += 1
    assert checked >= 4
