"""`emdFunction` / `emdModule` / `calc_emd` -- the reference's auction-EMD surface over the B200 cluster kernel.

Mirrors Density_aware_Chamfer_Distance/utils_v2/metrics/EMD/emd_module.py:39-91 (``emdModule()(input1, input2, eps, iters)
-> (dist [B, n], assignment [B, n] int32)``; only input1 receives a gradient) and ``calc_emd`` of
Density_aware_Chamfer_Distance/utils_v2/model_utils.py:72-77.  In U-RED this is the re-rank step after the Chamfer top-k
(engine/generate_pair.py:95-104: the 20 best sources by cd_m are re-scored with EMD); `rerank_emd` is that step.

Differences, all deliberate: one launch for the whole auction instead of seven per iteration; work on the caller's current
stream; no limit on the batch size (the reference asserts B <= 512) or n % 1024; a failing native call raises.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _native
from .dist_chamfer_3D import _require_cloud, _stream


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        lib = _native.load()
        _require_cloud("xyz1", xyz1)
        _require_cloud("xyz2", xyz2)
        B, n, _ = xyz1.shape
        if tuple(xyz2.shape) != (B, n, 3):
            raise AssertionError(f"EMD needs two clouds of equal size, got {tuple(xyz1.shape)} and {tuple(xyz2.shape)}")  # emd_module.py:46-47
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        dev = xyz1.device
        dist = torch.empty(B, n, device=dev, dtype=torch.float32)
        assignment = torch.empty(B, n, device=dev, dtype=torch.int32)
        ws_bytes = lib.ured_emd_workspace_bytes(B, n)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            rc = lib.ured_emd_forward(_native.ptr(xyz1), _native.ptr(xyz2), B, n, float(eps), int(iters),
                                      _native.ptr(dist), _native.ptr(assignment), _native.ptr(ws), ws_bytes, _stream(dev))
        _native.check(rc, "ured_emd_forward")
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, _gradidx):
        lib = _native.load()
        xyz1, xyz2, assignment = ctx.saved_tensors
        B, n, _ = xyz1.shape
        graddist = graddist.contiguous().float()
        gradxyz1 = torch.empty_like(xyz1)
        with torch.cuda.device(xyz1.device):
            rc = lib.ured_emd_backward(_native.ptr(xyz1), _native.ptr(xyz2), B, n, _native.ptr(graddist), _native.ptr(assignment),
                                       _native.ptr(gradxyz1), _stream(xyz1.device))
        _native.check(rc, "ured_emd_backward")
        return gradxyz1, torch.zeros_like(xyz2), None, None   # the reference returns an all-zero gradient for xyz2 (emd_module.py:81-84)


class emdModule(nn.Module):
    """emd_module.py:86-91."""

    def __init__(self):
        super(emdModule, self).__init__()

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1.float(), input2.float(), eps, iters)


def calc_emd(output, gt, eps=0.005, iterations=50):
    """model_utils.py:72-77: (mean over points of sqrt(dist) [B], dist [B, n])."""
    dist, _ = emdModule()(output, gt, eps, iterations)
    emd_out = torch.sqrt(dist).mean(1)
    return emd_out, dist


def rerank_emd(targets, library, ids, eps=0.005, iterations=50):
    """Re-rank retrieved candidates by EMD (engine/generate_pair.py:95-104: top-k by cd_m, then EMD on those k).

    targets [Q, n, 3], library [S, n, 3] (tensor), ids [Q, k] int (library indices, e.g. from `retrieve`).
    Returns (emd [Q, k] sorted ascending, ids [Q, k] in that order); ties keep the incoming (Chamfer) order.
    Follows compute_emd_loss2(p1=target, p2=source) -> calc_emd(output=p1, gt=p2) (engine/geometry_utils.py:84-86).
    """
    Q, k = ids.shape
    n = targets.shape[1]
    cand = library[ids.reshape(-1).long()]                       # [Q*k, n, 3]
    tgt = targets.repeat_interleave(k, dim=0)
    if cand.shape[1] != n:
        raise ValueError("EMD needs targets and library shapes of equal size")
    with torch.no_grad():
        emd, _ = calc_emd(tgt.float(), cand.float(), eps, iterations)
    emd = emd.view(Q, k)
    order = torch.sort(emd, dim=1, stable=True).indices
    return torch.gather(emd, 1, order), torch.gather(ids, 1, order.to(ids.device))
