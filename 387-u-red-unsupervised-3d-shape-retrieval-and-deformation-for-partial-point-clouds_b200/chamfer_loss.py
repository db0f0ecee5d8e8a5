"""U-RED's Chamfer loss callers over the B200 op (loss/chamfer_loss.py:5-30).

The reference imports ``Shape_Measure.distance.ChamferLoss`` -- a module that is neither vendored
nor pinned anywhere in the reference tree.  ``ChamferLoss`` below is the adaptor
for that call site under the documented assumption that ``ChamferLoss()(p1, p2)`` returns the
per-point SQUARED nearest-neighbour costs ``(cost1 [B,N], cost2 [B,M])``, i.e. exactly
``(dist1, dist2)`` of ``chamfer_3DDist``.  Parity at this boundary is unpinned by the reference.
"""
import torch
from torch import nn

from .dist_chamfer_3D import chamfer_3DDist


class ChamferLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self._cd = chamfer_3DDist()

    def forward(self, p1, p2):
        cost1, cost2, _, _ = self._cd(p1.float(), p2.float())
        return cost1, cost2


def chamfer_distance2(p1, p2):
    """loss/chamfer_loss.py:5-10 -- mean(cost1, 1) + mean(cost2, 1), one value per sample."""
    cost1, cost2 = ChamferLoss()(p1, p2)
    return cost1.mean(dim=1) + cost2.mean(dim=1)


def compute_cm_loss(source_p, target_p, target_part, mask=None, batch_reduction="mean"):
    """loss/chamfer_loss.py:13-30 -- full-shape and per-part Chamfer of deformed sources, as TWO batched calls.

    With ``mask`` ([B, parts]) sample ``bs`` uses its first ``mask[bs].sum() * 1024`` source points against the
    2048-point target, and part ``i`` uses source points [i*1024, (i+1)*1024) against the ragged
    ``target_part[bs][i]``.  Returns (mean full loss, mean part loss).  Without a mask it is
    ``chamfer_distance2(source_p, target_p)``.

    The reference loops over samples and parts in Python with one ``.item()`` sync and one B=1 Chamfer call per
    sample and per part; here the source lengths stay on the device and both the full-shape and the per-part
    losses are one ragged batched call each (``chamfer_ragged``).
    """
    if mask is None:
        return chamfer_distance2(source_p, target_p)
    from .model_utils import chamfer_ragged
    B = source_p.shape[0]
    dev = source_p.device
    src = source_p.float()
    # ---- full shapes: x = source (ragged length), gt = target ---------------------------------
    len_src = (mask.sum(1) * 1024).to(torch.int32)
    _, _, full, _, _, _, _ = chamfer_ragged(src, target_p.float(), len_x=len_src, alpha=0.0)
    # ---- parts: one pair per (sample, part) present in target_part ----------------------------
    n_parts = [len(tp) for tp in target_part]
    P = max(n_parts) if n_parts else 0
    if P == 0:
        return full.mean(), full.new_zeros(())
    if src.shape[1] < P * 1024:
        raise ValueError("source_p has fewer than 1024 points per target part")
    # Padded [B*P, m_max, 3] targets from the ragged parts with a fixed number of device operations -- one cat, one
    # scatter -- whatever B and P are (the row / column of every point is computed on the host from the shapes alone).
    import numpy as np
    sizes = np.zeros(B * P, dtype=np.int64)
    pieces = []
    for bs, tp in enumerate(target_part):
        for i, t in enumerate(tp):
            sizes[bs * P + i] = t.shape[0]
            if t.shape[0]:
                pieces.append(t)
    m_max = int(sizes.max())
    if m_max == 0:
        return full.mean(), full.new_zeros(())
    flat = torch.cat(pieces).to(device=dev, dtype=src.dtype)                        # [sum(sizes), 3]
    row_of = np.repeat(np.arange(B * P, dtype=np.int64), sizes)
    col_of = np.arange(int(sizes.sum()), dtype=np.int64) - np.repeat(np.cumsum(sizes) - sizes, sizes)
    where = torch.from_numpy(row_of * m_max + col_of).to(dev, non_blocking=True)    # one H2D copy
    tgt = src.new_zeros(B * P * m_max, 3).index_copy_(0, where, flat).view(B * P, m_max, 3)
    len_tgt = torch.from_numpy(sizes.astype(np.int32)).to(dev, non_blocking=True)
    parts_src = src[:, :P * 1024].reshape(B * P, 1024, 3)
    _, _, per_part, _, _, _, _ = chamfer_ragged(parts_src, tgt, len_gt=len_tgt, alpha=0.0)
    present = (len_tgt > 0).view(B, P).to(per_part.dtype)
    loss_part = (per_part.view(B, P) * present).sum(1) / present.sum(1)
    return full.mean(), loss_part.mean()
