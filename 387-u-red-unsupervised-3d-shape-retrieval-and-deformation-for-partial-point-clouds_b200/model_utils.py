"""`calc_dcd` / `calc_cd` -- the reference's metric functions over the fused B200 path.

Mirrors Density_aware_Chamfer_Distance/utils_v2/model_utils.py:13-70 (same names, arguments,
return ordering, dtypes).  Where the reference runs the native Chamfer op and then ~25 small
torch kernels (exp, scatter_add_, gather, pow, mean, four int64 casts), this runs
    pack -> nn_kernel (both directions) -> dcd_fwd_kernel
forward and one fused gradient pass backward.  Values agree with the reference to float32
round-off (the per-pair means are accumulated in float64 here); idx/dist are bit-identical.
"""
import torch
from torch.autograd import Function

from . import _native
from .dist_chamfer_3D import _len_arg, _require_cloud, _stream, chamfer_3DDist, nn_forward


def fscore(dist1, dist2, threshold=0.0001):
    """F-score of two clouds from their squared NN distances (metrics/CD/fscore.py:3-16).

    Returns (fscore, precision_1, precision_2), each [B]; 0/0 is reported as 0.
    """
    precision_1 = (dist1 < threshold).float().mean(dim=1)
    precision_2 = (dist2 < threshold).float().mean(dim=1)
    f = 2 * precision_1 * precision_2 / (precision_1 + precision_2)
    f = torch.where(torch.isnan(f), torch.zeros_like(f), f)
    return f, precision_1, precision_2


def torch_epilogue(dist1, dist2, idx1, idx2, alpha, n_lambda, frac_12, frac_21):
    """(loss, cd_p, cd_t) [B] through torch's own kernels, op for op as the reference's calc_dcd / calc_cd bodies
    (model_utils.py:26-45, 57-58), applied to given NN results of chamfer(gt, x).

    This is the audit / ``exact_ranking`` path: whatever torch's reduction kernels do for this shape and batch size,
    this does too, because it calls them (exp, scatter_add_, gather, pow, mean -- about 25 launches).  No gradient.
    """
    def density_term(dist, idx, bins_like, frac):
        hits = torch.zeros_like(bins_like).scatter_add_(1, idx.long(), torch.ones_like(idx))   # int32 histogram of the argmins
        weight = hits.gather(1, idx.long()).float() ** n_lambda
        weight = (weight + 1e-6) ** (-1) * frac
        return (1 - torch.exp(-dist * alpha) * weight).mean(dim=1)

    with torch.no_grad():
        loss = (density_term(dist1, idx1, idx2, frac_21) + density_term(dist2, idx2, idx1, frac_12)) / 2
        cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
        cd_t = dist1.mean(1) + dist2.mean(1)
    return loss, cd_p, cd_t


def fscore_fused(dist1, dist2, threshold=0.0001):
    """`fscore` as part of the epilogue kernel (one launch): (fscore, precision_1, precision_2), each [B]."""
    lib = _native.load()
    B, n1 = dist1.shape
    n2 = dist2.shape[1]
    dev = dist1.device
    out = torch.empty(3, B, device=dev, dtype=torch.float32)
    dummy = torch.empty(1, device=dev, dtype=torch.int32)   # idx is not read when no DCD term is requested
    with torch.cuda.device(dev):
        rc = lib.ured_dcd_forward_ex(_native.ptr(dist1.contiguous()), _native.ptr(dist2.contiguous()), _native.ptr(dummy), _native.ptr(dummy),
                                     B, n1, n2, 1, max(B, 1), None, None, 0.0, 1.0, 1.0, 1.0, 0,
                                     None, None, None, None, None, _native.ptr(out), float(threshold), _stream(dev))
    _native.check(rc, "ured_dcd_forward_ex")
    return out[0], out[1], out[2]


class _ChamferDCD(Function):
    """chamfer(gt, x) + the calc_cd/calc_dcd epilogue as one autograd node.

    forward(x, gt, alpha, n_lambda, frac_12, frac_21, len_x=None, len_gt=None, non_reg=False)
        -> (loss, cd_p, cd_t, dist1, dist2, idx1, idx2)   with cloud 1 = gt, cloud 2 = x
    len_x / len_gt: optional valid point counts per sample (ragged batches); the means then run over the valid
    points and the DCD fractions are rebuilt per sample from the lengths.
    """

    @staticmethod
    def forward(ctx, x, gt, alpha, n_lambda, frac_12, frac_21, len_x=None, len_gt=None, non_reg=False):
        lib = _native.load()
        B = x.shape[0]
        dev = x.device
        len1, len2 = _len_arg("len_gt", len_gt, B, dev), _len_arg("len_x", len_x, B, dev)
        dist1, dist2, idx1, idx2 = nn_forward(gt, x, len1=len1, len2=len2)
        n1 = dist1.shape[1]
        n2 = dist2.shape[1]
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        cd_p = torch.empty(B, device=dev, dtype=torch.float32)
        cd_t = torch.empty(B, device=dev, dtype=torch.float32)
        ew1 = torch.empty_like(dist1)
        ew2 = torch.empty_like(dist2)
        with torch.cuda.device(dev):
            rc = lib.ured_dcd_forward(_native.ptr(dist1), _native.ptr(dist2), _native.ptr(idx1), _native.ptr(idx2),
                                      B, n1, n2, 1, max(B, 1), _native.ptr(len1), _native.ptr(len2),
                                      float(alpha), float(n_lambda), float(frac_12), float(frac_21),
                                      _native.URED_FLAG_NON_REG if non_reg else 0,
                                      _native.ptr(loss), _native.ptr(cd_p), _native.ptr(cd_t),
                                      _native.ptr(ew1), _native.ptr(ew2), _stream(dev))
        _native.check(rc, "ured_dcd_forward")
        ctx.alpha = float(alpha)
        ctx.lens = (len1, len2)
        ctx.save_for_backward(x, gt, dist1, dist2, idx1, idx2, ew1, ew2)
        ctx.mark_non_differentiable(idx1, idx2)
        ctx.set_materialize_grads(False)
        return loss, cd_p, cd_t, dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, g_loss, g_cd_p, g_cd_t, g_dist1, g_dist2, _gi1, _gi2):
        lib = _native.load()
        x, gt, dist1, dist2, idx1, idx2, ew1, ew2 = ctx.saved_tensors
        B, n1 = dist1.shape
        n2 = dist2.shape[1]
        dev = x.device

        def prep(g):
            return None if g is None else g.contiguous().float()

        g_loss, g_cd_p, g_cd_t, g_dist1, g_dist2 = map(prep, (g_loss, g_cd_p, g_cd_t, g_dist1, g_dist2))
        grad_gt = torch.empty_like(gt)
        grad_x = torch.empty_like(x)
        with torch.cuda.device(dev):
            rc = lib.ured_dcd_backward(_native.ptr(gt), _native.ptr(x), B, n1, n2, 1, max(B, 1),
                                       _native.ptr(ctx.lens[0]), _native.ptr(ctx.lens[1]),
                                       _native.ptr(dist1), _native.ptr(dist2), _native.ptr(idx1), _native.ptr(idx2),
                                       _native.ptr(ew1), _native.ptr(ew2), ctx.alpha,
                                       _native.ptr(g_loss), _native.ptr(g_cd_p), _native.ptr(g_cd_t),
                                       _native.ptr(g_dist1), _native.ptr(g_dist2),
                                       _native.ptr(grad_gt), _native.ptr(grad_x), _stream(dev))
        _native.check(rc, "ured_dcd_backward")
        return grad_x, grad_gt, None, None, None, None, None, None, None


def _fused(x, gt, alpha, n_lambda, frac_12, frac_21, len_x=None, len_gt=None, non_reg=False):
    _require_cloud("x", x)
    _require_cloud("gt", gt)
    if x.shape[0] != gt.shape[0]:
        raise AssertionError(f"batch mismatch: {x.shape[0]} vs {gt.shape[0]}")  # model_utils.py:18 is an assert
    if x.shape[1] == 0 or gt.shape[1] == 0:
        raise ValueError("calc_cd / calc_dcd need non-empty clouds")
    return _ChamferDCD.apply(x.contiguous(), gt.contiguous(), alpha, n_lambda, frac_12, frac_21, len_x, len_gt, non_reg)


def calc_dcd(x, gt, alpha=1000, n_lambda=1, return_raw=False, non_reg=False):
    """Density-aware Chamfer distance (model_utils.py:13-51).

    Returns ``[loss, cd_p, cd_t]`` (each [B] float32), plus ``[dist1, dist2, idx1, idx2]`` when
    ``return_raw``; dist1/idx1 are [B, n_gt] (a gt point's nearest x point), dist2/idx2 [B, n_x].
    """
    x = x.float()
    gt = gt.float()
    n_x = x.shape[1]
    n_gt = gt.shape[1]
    assert x.shape[0] == gt.shape[0]
    if non_reg:
        frac_12 = max(1, n_x / n_gt)
        frac_21 = max(1, n_gt / n_x)
    else:
        frac_12 = n_x / n_gt
        frac_21 = n_gt / n_x
    loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = _fused(x, gt, alpha, n_lambda, frac_12, frac_21)
    res = [loss, cd_p, cd_t]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def calc_cd(output, gt, calc_f1=False, return_raw=False, normalize=False, separate=False):
    """Chamfer distance (model_utils.py:53-70): ``[cd_p, cd_t]`` (+ f1) (+ raw).

    ``normalize`` is accepted and ignored, exactly as in the reference.  ``separate`` returns the
    two directions unsummed, stacked [2, B]; that rarely used branch is computed with the
    reference's torch reductions on the kernel's dist1/dist2.
    """
    if separate:
        dist1, dist2, idx1, idx2 = chamfer_3DDist()(gt, output)
        res = [torch.stack([torch.sqrt(dist1).mean(1), torch.sqrt(dist2).mean(1)]),
               torch.stack([dist1.mean(1), dist2.mean(1)])]
    else:
        _loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = _fused(output, gt, 0.0, 1.0, 1.0, 1.0)
        res = [cd_p, cd_t]
    if calc_f1:
        f1, _, _ = fscore_fused(dist1.detach(), dist2.detach())   # fscore.py:3-16 inside the epilogue kernel
        res.append(f1)
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def chamfer_ragged(x, gt, len_x=None, len_gt=None, alpha=1000, n_lambda=1, non_reg=False):
    """Batched Chamfer/DCD over padded ragged clouds: sample b uses x[b, :len_x[b]] and gt[b, :len_gt[b]].

    Returns ``(loss, cd_p, cd_t, dist1, dist2, idx1, idx2)`` exactly as ``calc_dcd(..., return_raw=True)`` would for
    each sample on its own slices (entries past a sample's length are 0; a sample with an empty side gets zeros).
    The lengths may be device tensors -- no host synchronisation happens here.  This is the batched form of the
    per-sample loop in loss/chamfer_loss.py:13-30.
    """
    return _fused(x.float(), gt.float(), alpha, n_lambda, 1.0, 1.0, len_x, len_gt, non_reg)
