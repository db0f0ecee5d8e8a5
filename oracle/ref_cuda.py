"""The UNMODIFIED reference CUDA op (oracle/_ref, built by oracle/build.py) as a callable checker -- TEST INFRASTRUCTURE.

Used by tests/, tools/ and bench.py's non-product `reference_cuda_op` leg (the "beat the reference op recompiled for
sm_100a on the same box" bar of SURVEY.md 8(d) / BASELINE.md 4).  Nothing in the product package imports this.
"""
import torch

from . import build


def load():
    """The compiled reference module (chamfer_3D.forward / .backward of chamfer_cuda.cpp:17-33) or None."""
    return build.load_ref()


def forward(ref, xyz1, xyz2):
    """dist_chamfer_3D.py:28-46: zero-filled outputs, then the op."""
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1, d2 = torch.zeros(B, n, device=dev), torch.zeros(B, m, device=dev)
    i1, i2 = torch.zeros(B, n, device=dev, dtype=torch.int32), torch.zeros(B, m, device=dev, dtype=torch.int32)
    ref.forward(xyz1, xyz2, d1, d2, i1, i2)
    return d1, d2, i1, i2


def scores(ref, x, gt, alpha=1000, n_lambda=1):
    """(loss, cd_p, cd_t) [B] exactly as the reference computes them: its op + its torch ops (model_utils.py:13-58)."""
    from .torch_path import calc_dcd_oracle
    with torch.no_grad():
        loss, cd_p, cd_t = calc_dcd_oracle(x, gt, alpha=alpha, n_lambda=n_lambda, chamfer=lambda a, b: forward(ref, a, b))
    return loss, cd_p, cd_t


def calc_dcd_fwd_bwd(ref, x, gt, alpha=1000, n_lambda=1):
    """One training step of the reference on the GPU: op forward, torch-op DCD body, autograd to d loss / d dist,
    op backward (dist_chamfer_3D.py:49-64)."""
    n_x, n_gt = x.shape[1], gt.shape[1]
    d1, d2, i1, i2 = forward(ref, gt, x)
    d1.requires_grad_(); d2.requires_grad_()
    e1, e2 = torch.exp(-d1 * alpha), torch.exp(-d2 * alpha)
    c1 = torch.zeros_like(i2); c1.scatter_add_(1, i1.long(), torch.ones_like(i1))
    w1 = (c1.gather(1, i1.long()).float() ** n_lambda + 1e-6) ** (-1) * (n_gt / n_x)
    c2 = torch.zeros_like(i1); c2.scatter_add_(1, i2.long(), torch.ones_like(i2))
    w2 = (c2.gather(1, i2.long()).float() ** n_lambda + 1e-6) ** (-1) * (n_x / n_gt)
    loss = ((1 - e1 * w1).mean(1) + (1 - e2 * w2).mean(1)) / 2
    g1, g2 = torch.autograd.grad(loss.sum(), [d1, d2])
    gx1, gx2 = torch.zeros_like(gt), torch.zeros_like(x)
    ref.backward(gt, x, gx1, gx2, g1.contiguous(), g2.contiguous(), i1, i2)
    return loss


# ---- the reference's auction-EMD op (oracle/_ref/emd, built by oracle/build.py::build_ref_emd) ---------------------------
def load_emd():
    return build.load_ref_emd()


def emd_forward(ref_emd, xyz1, xyz2, eps, iters):
    """emd_module.py:41-68 around emd.forward: the reference's own buffer set-up, then its op.  n must be a multiple of 1024."""
    B, n, _ = xyz1.shape
    dev = xyz1.device
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
    dist = z(B, n)
    assignment = z(B, n, dt=torch.int32) - 1
    assignment_inv = z(B, n, dt=torch.int32) - 1
    price, bid, bid_increments, max_increments = z(B, n), z(B, n, dt=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dt=torch.int32), z(B * n, dt=torch.int32)
    unass_cnt, unass_cnt_sum, cnt_tmp = z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32)
    ref_emd.forward(xyz1.contiguous(), xyz2.contiguous(), dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
                    unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters)
    torch.cuda.synchronize()
    return dist, assignment
