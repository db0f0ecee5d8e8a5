"""GPU: bit-for-bit comparison with the UNMODIFIED reference CUDA op (oracle/_ref, built from
/root/reference by oracle/build.py; the prebuilt module travels to the GPU box).

This is what pins both the oracle and the product to the reference's own implementation:
  reference op == C oracle == B200 kernels (screen and exact) on identical inputs, including the
  BASELINE configs' full sizes, which the GPU reference finishes in milliseconds.
"""
import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_op(oracle):
    mod = oracle.build.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref not built (needs /root/reference in the authoring container)")
    return mod


def ref_forward(ref_op, a, b):
    """dist_chamfer_3D.py:28-47 around chamfer_3D.forward."""
    B, n, _ = a.shape
    m = b.shape[1]
    d1 = torch.zeros(B, n, device="cuda"); d2 = torch.zeros(B, m, device="cuda")
    i1 = torch.zeros(B, n, device="cuda", dtype=torch.int32); i2 = torch.zeros(B, m, device="cuda", dtype=torch.int32)
    assert ref_op.forward(a, b, d1, d2, i1, i2) == 1
    torch.cuda.synchronize()
    return d1, d2, i1, i2


CASES = [(4, 100, 200, "U"), (32, 2000, 1000, "U"), (32, 2048, 2048, "S"), (7, 1001, 517, "S"),
         (64, 2048, 2048, "U"), (2, 16384, 16384, "S"), (1, 16384, 16384, "U"), (3, 5000, 3000, "S")]


@pytest.mark.parametrize("B,N,M,kind", CASES)
def test_kernels_bit_exact_vs_reference_op(ured, ref_op, B, N, M, kind):
    a, b = make_clouds(0, B, N, kind).cuda(), make_clouds(1, B, M, kind).cuda()
    want = ref_forward(ref_op, a, b)
    for kernel in ("tensor", "fp32", "exact"):     # screening on the tensor cores / on the FP32 pipes / difference form on every pair
        got = ured.nn_forward(a, b, exact_only=kernel == "exact", fp32_screen=kernel == "fp32")
        for g, w, name in zip(got, want, ["dist1", "dist2", "idx1", "idx2"]):
            assert torch.equal(g, w), f"{name} differs from the reference op ({kernel} kernel): {(g != w).sum().item()} entries"


def test_c_oracle_bit_exact_vs_reference_op(oracle, ref_op):
    for (B, N, M, kind) in [(4, 100, 200, "U"), (2, 2000, 1000, "U"), (2, 2048, 2048, "S"), (1, 1500, 700, "S")]:
        a, b = make_clouds(5, B, N, kind), make_clouds(6, B, M, kind)
        want = [t.cpu().numpy() for t in ref_forward(ref_op, a.cuda(), b.cuda())]
        got = oracle.c.chamfer_forward(a.numpy(), b.numpy())
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    g = torch.Generator().manual_seed(7)
    la, lb = torch.randint(0, 4, (2, 700, 3), generator=g).float(), torch.randint(0, 4, (2, 900, 3), generator=g).float()
    want = [t.cpu().numpy() for t in ref_forward(ref_op, la.cuda(), lb.cuda())]
    for g_, w in zip(oracle.c.chamfer_forward(la.numpy(), lb.numpy()), want):
        assert np.array_equal(g_, w)


def test_backward_vs_reference_op(ured, ref_op):
    """Gradients: the reference's atomics are order-dependent, so 1e-5 relative (north_star)."""
    B, N, M = 8, 2048, 2048
    a, b = make_clouds(0, B, N, "S").cuda(), make_clouds(1, B, M, "S").cuda()
    g = torch.Generator().manual_seed(2)
    w1, w2 = torch.randn(B, N, generator=g).cuda(), torch.randn(B, M, generator=g).cuda()
    d1, d2, i1, i2 = ref_forward(ref_op, a, b)
    g1, g2 = torch.zeros_like(a), torch.zeros_like(b)
    assert ref_op.backward(a, b, g1, g2, w1, w2, i1, i2) == 1
    xa, xb = a.clone().requires_grad_(), b.clone().requires_grad_()
    e1, e2, _, _ = ured.chamfer_3DDist()(xa, xb)
    ((e1 * w1).sum() + (e2 * w2).sum()).backward()
    for got, want in [(xa.grad, g1), (xb.grad, g2)]:
        assert ((got - want).abs().max() / want.abs().max()).item() < 1e-5


def test_dcd_vs_reference_torch_ops_on_gpu(ured, ref_op):
    """calc_dcd's torch-op body (model_utils.py:13-51) run on the GPU over the reference op's outputs."""
    B, n_x, n_gt, alpha, lam = 16, 2048, 2048, 1000, 1
    x, gt = make_clouds(3, B, n_x, "S").cuda(), (make_clouds(4, B, n_gt, "S") * 0.95).cuda()
    dist1, dist2, idx1, idx2 = ref_forward(ref_op, gt, x)
    cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
    cd_t = dist1.mean(1) + dist2.mean(1)
    e1, e2 = torch.exp(-dist1 * alpha), torch.exp(-dist2 * alpha)
    c1 = torch.zeros_like(idx2); c1.scatter_add_(1, idx1.long(), torch.ones_like(idx1))
    w1 = (c1.gather(1, idx1.long()).float() ** lam + 1e-6) ** (-1) * (n_gt / n_x)
    c2 = torch.zeros_like(idx1); c2.scatter_add_(1, idx2.long(), torch.ones_like(idx2))
    w2 = (c2.gather(1, idx2.long()).float() ** lam + 1e-6) ** (-1) * (n_x / n_gt)
    loss = ((1 - e1 * w1).mean(1) + (1 - e2 * w2).mean(1)) / 2
    got = ured.calc_dcd(x, gt, alpha=alpha, n_lambda=lam)
    for g_, w_ in zip(got, [loss, cd_p, cd_t]):
        assert torch.allclose(g_, w_, rtol=1e-5, atol=0)
    # retrieval ranking by the reference's scores == ranking by ours
    assert torch.equal(torch.sort(cd_t, stable=True).indices.int(), ured.topk_smallest(got[2], B)[1])
