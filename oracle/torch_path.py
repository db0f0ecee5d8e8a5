"""Torch-on-CPU restatements -- TEST INFRASTRUCTURE ONLY.

dist_chamfer_cpu   <- chamfer_python.distChamfer      Density_aware_Chamfer_Distance/utils_v2/metrics/CD/chamfer_python.py:18-39
                      ("the repo's torch CPU path" of BASELINE.json config 0: float64 expansion form,
                      full [B,N,M] matrix, four min passes); also the CPU baseline that bench.py times.
OracleChamfer      <- chamfer_3DFunction               .../chamfer3D/dist_chamfer_3D.py:26-64, with the C oracle as the native op
calc_cd_oracle     <- calc_cd                          Density_aware_Chamfer_Distance/utils_v2/model_utils.py:53-70
calc_dcd_oracle    <- calc_dcd                         Density_aware_Chamfer_Distance/utils_v2/model_utils.py:13-51
The DCD functions take the Chamfer callable as an argument because the reference hard-wires the
CUDA op (model_utils.py:55); every torch op and rounding point is kept in the reference's order.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import chamfer_oracle


def dist_chamfer_cpu(a, b):
    """chamfer_python.py:18-39 -- returns (dist a->b, dist b->a, idx a->b, idx b->a)."""
    x = a.double()
    y = b.double()
    sq_x = (x * x).sum(-1)                      # [B, Nx]
    sq_y = (y * y).sum(-1)                      # [B, Ny]
    cross = torch.bmm(x, y.transpose(1, 2))     # [B, Nx, Ny]
    pair = sq_x.unsqueeze(2) + sq_y.unsqueeze(1) - 2 * cross
    # the reference evaluates min over each axis twice (once for values, once for indices)
    d_ab = torch.min(pair, 2)[0].float()
    d_ba = torch.min(pair, 1)[0].float()
    i_ab = torch.min(pair, 2)[1].int()
    i_ba = torch.min(pair, 1)[1].int()
    return d_ab, d_ba, i_ab, i_ba


class OracleChamfer(Function):
    """Autograd node around the C oracle (CPU tensors)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = chamfer_oracle.chamfer_forward(xyz1.detach().numpy(), xyz2.detach().numpy())
        d1, d2, i1, i2 = map(torch.from_numpy, (d1, d2, i1, i2))
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, g1, g2, _gi1, _gi2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        z = lambda g, ref: np.zeros(tuple(ref.shape), np.float32) if g is None else g.contiguous().numpy()
        gx1, gx2 = chamfer_oracle.chamfer_backward(xyz1.detach().numpy(), xyz2.detach().numpy(),
                                                   z(g1, i1), z(g2, i2), i1.numpy(), i2.numpy())
        return torch.from_numpy(gx1), torch.from_numpy(gx2)


def oracle_cd(xyz1, xyz2):
    return OracleChamfer.apply(xyz1.contiguous(), xyz2.contiguous())


def calc_cd_oracle(output, gt, chamfer=oracle_cd, return_raw=False):
    """model_utils.py:53-70 (default branch)."""
    dist1, dist2, idx1, idx2 = chamfer(gt, output)  # argument swap, :56
    cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
    cd_t = dist1.mean(1) + dist2.mean(1)
    res = [cd_p, cd_t]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def calc_dcd_oracle(x, gt, alpha=1000, n_lambda=1, return_raw=False, non_reg=False, chamfer=oracle_cd):
    """model_utils.py:13-51."""
    x = x.float()
    gt = gt.float()
    n_x, n_gt = x.shape[1], gt.shape[1]
    assert x.shape[0] == gt.shape[0]
    if non_reg:
        frac_12, frac_21 = max(1, n_x / n_gt), max(1, n_gt / n_x)
    else:
        frac_12, frac_21 = n_x / n_gt, n_gt / n_x
    cd_p, cd_t, dist1, dist2, idx1, idx2 = calc_cd_oracle(x, gt, chamfer=chamfer, return_raw=True)
    exp_dist1, exp_dist2 = torch.exp(-dist1 * alpha), torch.exp(-dist2 * alpha)

    count1 = torch.zeros_like(idx2)
    count1.scatter_add_(1, idx1.long(), torch.ones_like(idx1))
    weight1 = count1.gather(1, idx1.long()).float().detach() ** n_lambda
    weight1 = (weight1 + 1e-6) ** (-1) * frac_21
    loss1 = (1 - exp_dist1 * weight1).mean(dim=1)

    count2 = torch.zeros_like(idx1)
    count2.scatter_add_(1, idx2.long(), torch.ones_like(idx2))
    weight2 = count2.gather(1, idx2.long()).float().detach() ** n_lambda
    weight2 = (weight2 + 1e-6) ** (-1) * frac_12
    loss2 = (1 - exp_dist2 * weight2).mean(dim=1)

    loss = (loss1 + loss2) / 2
    res = [loss, cd_p, cd_t]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def topk_oracle(scores, k):
    """Ascending (score, index) order: torch.sort(stable=True) on the scores (ranking)."""
    s, i = torch.sort(scores, dim=-1, stable=True)
    return s[..., :k], i[..., :k].int()
