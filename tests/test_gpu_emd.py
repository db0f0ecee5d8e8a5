"""GPU: the cluster-based auction-EMD kernel (ured_emd_forward / emdModule / calc_emd / rerank_emd).

Checked against (1) the CPU restatement oracle/emd_oracle.c on any size, bit for bit, and (2) the reference's own op
(emd.cpp + emd_cuda.cu compiled unmodified for sm_100a into oracle/_ref/emd), bit for bit whenever the reference's result is
well defined: its GetMax lets the LAST store win among bidders within 1e-6 of an object's best bid, so two runs of the
reference can differ; the oracle reports such tie events and the comparison with the reference op is exact only on tie-free
auctions (the usual case), while kernel == oracle is demanded always (both take the lowest point index).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def clouds(seed, B, n):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, n, 3, generator=g), torch.rand(B, n, 3, generator=g)   # coordinates in [0, 1], as the reference expects


@pytest.fixture(scope="module")
def emd_oracle():
    from oracle import emd_oracle as eo
    return eo


@pytest.mark.parametrize("B,n,eps,iters", [(2, 256, 0.005, 50), (3, 777, 0.005, 50), (1, 1024, 0.002, 200), (2, 2048, 0.005, 50),
                                           (1, 64, 0.01, 1), (2, 100, 0.005, 2), (1, 5000, 0.005, 10)])
def test_emd_kernel_equals_cpu_oracle(ured, emd_oracle, B, n, eps, iters):
    a, b = clouds(n + iters, B, n)
    dist, assignment = ured.emdModule()(a.cuda(), b.cuda(), eps, iters)
    wd, wa, _ = emd_oracle.emd_forward(a.numpy(), b.numpy(), eps, iters)
    assert np.array_equal(assignment.cpu().numpy(), wa), f"{(assignment.cpu().numpy() != wa).sum()} assignments differ"
    assert np.array_equal(dist.cpu().numpy(), wd)


def test_emd_kernel_equals_reference_op(ured, oracle, emd_oracle):
    from oracle import ref_cuda
    ref = ref_cuda.load_emd()
    if ref is None:
        pytest.skip("oracle/_ref/emd not built (needs /root/reference in the authoring container)")
    checked = 0
    for seed, (B, n, eps, iters) in enumerate([(4, 1024, 0.005, 50), (2, 2048, 0.005, 50), (2, 2048, 0.002, 300), (20, 2048, 0.005, 50)]):
        a, b = clouds(500 + seed, B, n)
        rd, ra = ref_cuda.emd_forward(ref, a.cuda(), b.cuda(), eps, iters)
        dist, assignment = ured.emdModule()(a.cuda(), b.cuda(), eps, iters)
        for s in range(min(B, 4)):     # the tie report comes from the CPU oracle, one pair at a time
            _, _, ties = emd_oracle.emd_forward(a[s:s + 1].numpy(), b[s:s + 1].numpy(), eps, iters)
            if ties == 0:
                assert torch.equal(assignment[s], ra[s]), f"pair {s} of case {seed}: assignment differs from the reference op"
                assert torch.equal(dist[s], rd[s])
                checked += 1
        # aggregate agreement even where a tie made the reference's own result run-dependent
        e_ref, e_got = torch.sqrt(rd).mean(1), torch.sqrt(dist).mean(1)
        assert torch.allclose(e_got, e_ref, rtol=2e-2)
    assert checked >= 8


def test_calc_emd_backward_and_rerank(ured, emd_oracle):
    B, n = 3, 512
    a, b = clouds(77, B, n)
    x = a.cuda().requires_grad_()
    emd_out, dist = ured.calc_emd(x, b.cuda(), eps=0.005, iterations=50)
    wd, wa, _ = emd_oracle.emd_forward(a.numpy(), b.numpy(), 0.005, 50)
    assert np.allclose(emd_out.detach().cpu().numpy(), np.sqrt(wd).mean(1), rtol=1e-6)
    w = torch.linspace(0.5, 2.0, B * n).view(B, n).cuda()
    (dist * w).sum().backward()
    matched = torch.gather(b.cuda(), 1, torch.from_numpy(wa).long().cuda().unsqueeze(-1).expand(-1, -1, 3))
    want = 2 * w.unsqueeze(-1) * (a.cuda() - matched)             # emd_cuda.cu:279-300; xyz2 gets no gradient
    assert torch.allclose(x.grad, want, rtol=1e-6, atol=1e-7)
    # re-rank: 2 targets, 6 library shapes, the Chamfer top-4 re-ordered by EMD
    lib = torch.rand(6, n, 3, generator=torch.Generator().manual_seed(5)).cuda()
    tg = (lib[[4, 1]] + 0.01 * torch.randn(2, n, 3, generator=torch.Generator().manual_seed(6)).cuda()).clamp(0, 1)
    _, ids = ured.retrieve(tg, lib, k=4)
    emd_sorted, ids_sorted = ured.rerank_emd(tg, lib, ids, eps=0.005, iterations=50)
    assert ids_sorted[:, 0].tolist() == [4, 1] and (emd_sorted[:, 1:] >= emd_sorted[:, :-1]).all()
    for q in range(2):
        for c in range(4):
            e, _ = ured.calc_emd(tg[q:q + 1], lib[ids_sorted[q, c].long()].unsqueeze(0))
            assert torch.allclose(e[0], emd_sorted[q, c], rtol=1e-6)   # (torch reduces a 1-row and an 8-row batch in different orders)


def test_emd_argument_errors(ured):
    with pytest.raises(AssertionError):
        ured.emdModule()(torch.rand(1, 8, 3).cuda(), torch.rand(1, 9, 3).cuda(), 0.005, 10)
    with pytest.raises(RuntimeError, match="GPU tensors only"):
        ured.emdModule()(torch.rand(1, 8, 3), torch.rand(1, 8, 3), 0.005, 10)
