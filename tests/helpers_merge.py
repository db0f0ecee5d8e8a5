"""Test helper: CPU reference of the (score, id) merge used by the exchange tests."""
import torch


def cpu_merge(scores, ids, k):
    """Reference merge for the CPU tests of the exchange logic: ascending (score, id), ids < 0 are padding.
    (The product's merge_topk is the native kernel and takes GPU tensors only.)"""
    scores = torch.where(ids < 0, torch.full_like(scores, float("inf")), scores)
    big = torch.iinfo(ids.dtype).max
    order = torch.sort(torch.where(ids < 0, torch.full_like(ids, big), ids), dim=1, stable=True).indices
    s = torch.gather(scores, 1, order)
    i = torch.gather(ids, 1, order)
    order = torch.sort(s, dim=1, stable=True).indices
    return torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
