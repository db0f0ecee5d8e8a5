"""GPU: GraphedDCD (CUDA-graph replay of the calc_dcd training step) equals the eager path, values and gradients."""
import time

import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


def test_graphed_dcd_matches_eager(ured):
    B, N = 32, 2048
    dcd = ured.GraphedDCD(B, N, N, alpha=1000, n_lambda=1)
    for seed in (200, 202, 200):
        x0, gt0 = make_clouds(seed, B, N, "S").cuda(), (make_clouds(seed + 1, B, N, "S") * 0.95).cuda()
        w = torch.linspace(0.5, 1.5, B, device="cuda")
        x, gt = x0.clone().requires_grad_(), gt0.clone().requires_grad_()
        loss, cd_p, cd_t = dcd(x, gt)
        (loss * w).sum().backward()
        xe, gte = x0.clone().requires_grad_(), gt0.clone().requires_grad_()
        el, ep, et = ured.calc_dcd(xe, gte, alpha=1000, n_lambda=1)
        (el * w).sum().backward()
        assert torch.equal(loss, el) and torch.equal(cd_p, ep) and torch.equal(cd_t, et)
        for got, want in [(x.grad, xe.grad), (gt.grad, gte.grad)]:
            assert ((got - want).abs().max() / want.abs().max()).item() < 1e-5
    with pytest.raises(ValueError):
        dcd(x0[:4], gt0[:4])

    def bench(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3

    def eager():
        a, b = x0.detach().requires_grad_(), gt0.detach().requires_grad_()
        ured.calc_dcd(a, b)[0].sum().backward()

    def graphed():
        a, b = x0.detach().requires_grad_(), gt0.detach().requires_grad_()
        dcd(a, b)[0].sum().backward()
    print(f"\ncfg1 step: eager {bench(eager):.3f} ms, graphed {bench(graphed):.3f} ms")


def test_graphed_whole_step_replay(ured):
    """GraphedDCD.forward_backward: loss and d sum(loss)/d clouds of one replay == the eager autograd step, bit for bit
    (the backward kernel is deterministic), and stays correct when inputs change between replays."""
    B, N = 16, 2048
    dcd = ured.GraphedDCD(B, N, N, alpha=200, n_lambda=0.5)
    for seed in (300, 304, 300):
        x0, gt0 = make_clouds(seed, B, N, "S").cuda(), (make_clouds(seed + 1, B, N, "S") * 0.9).cuda()
        loss, cd_p, cd_t, gx, ggt = dcd.forward_backward(x0, gt0)
        xe, gte = x0.clone().requires_grad_(), gt0.clone().requires_grad_()
        el, ep, et = ured.calc_dcd(xe, gte, alpha=200, n_lambda=0.5)
        el.sum().backward()
        assert torch.equal(loss, el) and torch.equal(cd_p, ep) and torch.equal(cd_t, et)
        assert torch.equal(gx, xe.grad) and torch.equal(ggt, gte.grad)
