"""CPU: pin the oracle against the reference-generated golden vectors (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g_py():
    return np.load(os.path.join(GOLD, "chamfer_ref_python.npz"))


@pytest.fixture(scope="module")
def g_dcd():
    return np.load(os.path.join(GOLD, "dcd_ref_model_utils.npz"))


@pytest.mark.parametrize("case", ["unit_test", "timing", "shape", "chair"])
def test_oracle_matches_reference_python_chamfer(oracle, g_py, case):
    """The reference's own criterion (unit_test.py:23-33): MSE < 1e-8 and idx exactly equal."""
    d1, d2, i1, i2 = oracle.c.chamfer_forward(g_py[f"{case}_xyz1"], g_py[f"{case}_xyz2"])
    mse = np.mean((d1 - g_py[f"{case}_dist1"]) ** 2) + np.mean((d2 - g_py[f"{case}_dist2"]) ** 2)
    assert mse < 1e-8
    assert np.array_equal(i1, g_py[f"{case}_idx1"])
    assert np.array_equal(i2, g_py[f"{case}_idx2"])
    # float32 difference form vs float64 expansion form: agree to float32 round-off of the coordinates
    assert np.allclose(d1, g_py[f"{case}_dist1"], rtol=1e-4, atol=1e-6)


def test_torch_cpu_port_matches_reference_python_chamfer(oracle, g_py):
    """oracle.torch_path.dist_chamfer_cpu is the CPU baseline bench.py times: same bits as the reference's."""
    for case in ["unit_test", "shape"]:
        d1, d2, i1, i2 = oracle.t.dist_chamfer_cpu(torch.from_numpy(g_py[f"{case}_xyz1"]), torch.from_numpy(g_py[f"{case}_xyz2"]))
        assert np.array_equal(i1.numpy(), g_py[f"{case}_idx1"]) and np.array_equal(i2.numpy(), g_py[f"{case}_idx2"])
        assert np.allclose(d1.numpy(), g_py[f"{case}_dist1"], rtol=1e-6, atol=1e-9)
        assert np.allclose(d2.numpy(), g_py[f"{case}_dist2"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("case", ["default", "pcn", "vrc_nonreg", "lambda2"])
def test_dcd_restatement_matches_reference_model_utils(oracle, g_dcd, case):
    """calc_dcd_oracle / calc_cd_oracle restate model_utils.py:13-70; goldens came from the unmodified file."""
    alpha, lam, non_reg = g_dcd[f"{case}_meta"]
    lam = int(lam) if float(lam).is_integer() else float(lam)
    x = torch.from_numpy(g_dcd[f"{case}_x"]).requires_grad_()
    gt = torch.from_numpy(g_dcd[f"{case}_gt"]).requires_grad_()
    w = torch.from_numpy(g_dcd[f"{case}_w"])
    loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = oracle.t.calc_dcd_oracle(x, gt, alpha=alpha, n_lambda=lam,
                                                                        return_raw=True, non_reg=bool(non_reg))
    assert np.array_equal(idx1.numpy(), g_dcd[f"{case}_idx1"]) and np.array_equal(idx2.numpy(), g_dcd[f"{case}_idx2"])
    assert np.array_equal(dist1.detach().numpy(), g_dcd[f"{case}_dist1"])
    for name, val in [("loss", loss), ("cd_p", cd_p), ("cd_t", cd_t)]:
        assert np.allclose(val.detach().numpy(), g_dcd[f"{case}_{name}"], rtol=1e-6, atol=0), name
    (loss * w).sum().backward()
    assert np.allclose(x.grad.numpy(), g_dcd[f"{case}_g_x_loss"], rtol=1e-5, atol=1e-9)
    assert np.allclose(gt.grad.numpy(), g_dcd[f"{case}_g_gt_loss"], rtol=1e-5, atol=1e-9)


def test_oracle_tie_and_tile_rules(oracle):
    """Lowest index wins on exact ties, also across the reference's 512-candidate tiles (chamfer3D.cu:36,126)."""
    q = np.zeros((1, 1, 3), np.float32)
    c = np.ones((1, 1500, 3), np.float32)
    c[0, [7, 600, 1400]] = 0.5  # three equidistant nearest candidates in three different tiles
    d1, _, i1, _ = oracle.c.chamfer_forward(q, c)
    assert i1[0, 0] == 7 and d1[0, 0] == np.float32(0.75)
    c[0, 3] = 0.5
    assert oracle.c.chamfer_forward(q, c)[2][0, 0] == 3


def test_oracle_backward_matches_autograd_of_python_chamfer(oracle):
    """Gradient of sum(w1*dist1)+sum(w2*dist2): C oracle (chamfer3D.cu:155-195) vs autograd through the float64 path."""
    from conftest import make_clouds
    a, b = make_clouds(3, 2, 150, "S"), make_clouds(4, 2, 220, "S")
    g = torch.Generator().manual_seed(5)
    w1, w2 = torch.randn(2, 150, generator=g), torch.randn(2, 220, generator=g)
    d1, d2, i1, i2 = oracle.c.chamfer_forward(a.numpy(), b.numpy())
    g1, g2 = oracle.c.chamfer_backward(a.numpy(), b.numpy(), w1.numpy(), w2.numpy(), i1, i2)
    a64, b64 = a.double().requires_grad_(), b.double().requires_grad_()
    pair = ((a64.unsqueeze(2) - b64.unsqueeze(1)) ** 2).sum(-1)
    (pair.min(2)[0] * w1.double()).sum().add((pair.min(1)[0] * w2.double()).sum()).backward()
    assert np.allclose(g1, a64.grad.numpy(), rtol=1e-4, atol=1e-6)
    assert np.allclose(g2, b64.grad.numpy(), rtol=1e-4, atol=1e-6)
    h1, h2 = oracle.c.chamfer_backward_f64(a.numpy(), b.numpy(), w1.numpy(), w2.numpy(), i1, i2)
    assert np.allclose(g1, h1, rtol=1e-5, atol=1e-7) and np.allclose(g2, h2, rtol=1e-5, atol=1e-7)


CUDA_CASES = ["unit_test", "timing", "ragged_tail", "lattice", "chair"]


@pytest.mark.parametrize("case", CUDA_CASES)
def test_oracle_bit_exact_vs_reference_cuda_op_golden(oracle, case):
    """tests/golden/chamfer_ref_cuda_b200.npz holds outputs of the UNMODIFIED reference CUDA op run on a B200
    (tests/golden/make_golden_gpu.py): the C oracle must reproduce dist/idx bit for bit, gradients to 1e-5."""
    g = np.load(os.path.join(GOLD, "chamfer_ref_cuda_b200.npz"))
    a, b = g[f"{case}_xyz1"], g[f"{case}_xyz2"]
    d1, d2, i1, i2 = oracle.c.chamfer_forward(a, b)
    assert np.array_equal(d1.view(np.uint32), g[f"{case}_dist1"].view(np.uint32))
    assert np.array_equal(d2.view(np.uint32), g[f"{case}_dist2"].view(np.uint32))
    assert np.array_equal(i1, g[f"{case}_idx1"]) and np.array_equal(i2, g[f"{case}_idx2"])
    r1, r2 = oracle.c.chamfer_backward_f64(a, b, g[f"{case}_w1"], g[f"{case}_w2"], i1, i2)
    for got, want in [(g[f"{case}_grad1"], r1), (g[f"{case}_grad2"], r2)]:
        assert np.abs(got - want).max() / (np.abs(want).max() + 1e-30) < 1e-5
