"""K=1 nearest neighbours with gathered points -- the `pytorch3d.ops.knn_points(p1, p2, K=1, return_nn=True)`
call of `residual_retrieval_loss` (loss/basic_loss.py:249-265) on the Chamfer NN kernel (one direction only).

pytorch3d is neither vendored nor installed in the reference tree, so parity at this call site is unpinned by the
reference; the contract here is: squared L2 distances, lowest index on ties (the Chamfer op's rule), `nn` gathered
from p2 and differentiable in p2, `dists` differentiable in both clouds.
"""
import torch
from torch.autograd import Function

from . import _native
from .dist_chamfer_3D import _len_arg, _require_cloud, _stream


class _NearestOne(Function):
    @staticmethod
    def forward(ctx, p1, p2, len2):
        lib = _native.load()
        B, n1, _ = p1.shape
        n2 = p2.shape[1]
        dev = p1.device
        dist1 = torch.empty(B, n1, device=dev, dtype=torch.float32)
        idx1 = torch.empty(B, n1, device=dev, dtype=torch.int32)
        pk2 = torch.empty(lib.ured_packed_bytes(B, n2), device=dev, dtype=torch.uint8)
        sb = lib.ured_nn_scratch_bytes(B, n1, n2)
        scratch = torch.empty(sb, device=dev, dtype=torch.uint8) if sb else None
        with torch.cuda.device(dev):
            _native.check(lib.ured_pack_clouds(_native.ptr(p2), B, n2, _native.ptr(len2), _native.ptr(pk2), _stream(dev)), "ured_pack_clouds")
            rc = lib.ured_nn_packed(_native.ptr(p1), None, n1,  # the image of p1 is not needed in one-direction mode
                                    _native.ptr(p2), _native.ptr(pk2), n2,
                                    B, 1, max(B, 1), None, _native.ptr(len2),
                                    _native.ptr(dist1), None, _native.ptr(idx1), None,
                                    _native.ptr(scratch), sb, _native.URED_FLAG_ONE_DIRECTION, _stream(dev))
        _native.check(rc, "ured_nn_packed")
        ctx.save_for_backward(p1, p2, idx1)
        ctx.len2 = len2
        ctx.mark_non_differentiable(idx1)
        ctx.set_materialize_grads(False)
        return dist1, idx1

    @staticmethod
    def backward(ctx, g_dist, _g_idx):
        p1, p2, idx1 = ctx.saved_tensors
        if g_dist is None:
            return None, None, None
        lib = _native.load()
        B, n1, _ = p1.shape
        n2 = p2.shape[1]
        g_dist = g_dist.contiguous().float()
        g1, g2 = torch.empty_like(p1), torch.empty_like(p2)
        with torch.cuda.device(p1.device):   # cloud-2 points have no term of their own: only the scatter side runs for them
            rc = lib.ured_nn_backward_one_direction(_native.ptr(p1), _native.ptr(p2), B, n1, n2, _native.ptr(ctx.len2), _native.ptr(g_dist),
                                                    _native.ptr(idx1), _native.ptr(g1), _native.ptr(g2), _stream(p1.device))
        _native.check(rc, "ured_nn_backward_one_direction")
        return g1, g2, None


def knn1_points(p1, p2, lengths2=None, return_nn=True):
    """(dists [B,N,1], idx [B,N,1] int64, nn [B,N,1,3] or None): nearest point of p2[b, :lengths2[b]] for every p1 point."""
    p1, p2 = p1.float().contiguous(), p2.float().contiguous()
    _require_cloud("p1", p1)
    _require_cloud("p2", p2)
    len2 = _len_arg("lengths2", lengths2, p1.shape[0], p1.device)
    dist, idx = _NearestOne.apply(p1, p2, len2)
    idx64 = idx.long()
    nn = None
    if return_nn:
        nn = torch.gather(p2, 1, idx64.unsqueeze(-1).expand(-1, -1, 3)).unsqueeze(2)
    return dist.unsqueeze(-1), idx64.unsqueeze(-1), nn


def residual_retrieval_loss(x, x_source, residuals, mask_part=None):
    """loss/basic_loss.py:249-265 without the per-sample loop and its `.item()` syncs.

    x [bs, n, 3] input points, x_source [bs, parts*1024, 3] deformed source (first mask_part.sum(1)*1024 points valid),
    residuals [bs, n, 3].  Returns (residual_loss, residual_loss_reg).
    """
    lengths = None if mask_part is None else (mask_part.sum(1) * 1024).to(torch.int32)
    _, _, nn = knn1_points(x, x_source, lengths2=lengths, return_nn=True)
    res_nn = x + residuals - nn.squeeze(2)
    residual_loss = torch.mean(torch.sum(torch.abs(res_nn), dim=-1))
    residual_loss_reg = torch.mean(torch.sum(torch.abs(residuals), dim=-1))
    return residual_loss, residual_loss_reg
