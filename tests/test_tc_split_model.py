"""CPU: the algebra behind the tensor-core screening pass (oracle/tc_split_model.py restates the kernel's operand rows).

(1) three bf16 pieces reproduce an fp32 value exactly; (2) the 27 piece products of a (query, candidate) pair reproduce
W_c - 2 q.c to ~2^-30 of the bound's scale S^2 -- so the only inexact step of the GPU kernel's screen is the tensor core's
fp32 accumulation, which tests/test_gpu_tensor_screen.py measures on the device; (3) the screen therefore ranks chunks like
the exact distances do: the argmin of the modelled scores is the reference argmin on random clouds.
"""
import numpy as np
import pytest

from conftest import make_clouds


@pytest.fixture(scope="module")
def model(oracle):
    from oracle import tc_split_model
    return tc_split_model


def test_three_bf16_pieces_are_exact(model):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000), rng.standard_normal(2000) * 1e6, rng.standard_normal(2000) * 1e-6,
                        rng.uniform(-1, 1, 2000) + 1000.0, [0.0, 1.0, -1.0, 3.0e38, 1e-30, 16777217.0, 0.1, 1.0 / 3.0]]).astype(np.float32)
    p1, p2, p3 = model.split3(x)
    for p in (p1, p2, p3):
        assert np.array_equal(p.view(np.uint32) & 0xFFFF, np.zeros(p.shape, np.uint32)), "a piece is not a bf16 value"
    assert np.array_equal((p1.astype(np.float64) + p2.astype(np.float64) + p3.astype(np.float64)).astype(np.float32), x)
    assert np.array_equal(p1.astype(np.float64) + p2.astype(np.float64) + p3.astype(np.float64), x.astype(np.float64)), "not exact"


@pytest.mark.parametrize("shift,scale_q,scale_c", [(0, 1, 1), (1000, 1, 1), (-37.5, 1, 1), (0, 1e3, 1), (0, 1, 1e3), (0, 1e-3, 1e-3), (3, 1e-2, 1e-2)])
def test_piece_products_reproduce_the_screening_score(model, shift, scale_q, scale_c):
    rng = np.random.default_rng(1)
    q = (rng.uniform(-1, 1, (128, 3)) * scale_q + shift).astype(np.float32)
    c = (rng.uniform(-1, 1, (256, 3)) * scale_c + shift).astype(np.float32)
    s, w = model.screen_scores(q, c)
    exact = w.astype(np.float64)[None, :] - 2.0 * (q.astype(np.float64) @ c.astype(np.float64).T)
    S = np.linalg.norm(q.astype(np.float64), axis=1)[:, None] + np.sqrt(w.astype(np.float64)).max()
    rel = np.abs(s - exact) / (S * S)
    assert rel.max() < 2.0 ** -29, f"dropped terms too large: 2^{np.log2(rel.max()):.1f}"


def test_modelled_screen_finds_the_reference_argmin(model, oracle):
    a, b = make_clouds(0, 1, 300, "S")[0].numpy(), make_clouds(1, 1, 500, "S")[0].numpy()
    want = oracle.c.chamfer_forward(a[None], b[None])
    s, _ = model.screen_scores(a, b)
    qn = (a.astype(np.float64) ** 2).sum(1)[:, None]
    d_model = s + qn
    assert np.array_equal(d_model.argmin(1), want[2][0])
    assert np.allclose(d_model.min(1), want[0][0], rtol=1e-5, atol=1e-7)
