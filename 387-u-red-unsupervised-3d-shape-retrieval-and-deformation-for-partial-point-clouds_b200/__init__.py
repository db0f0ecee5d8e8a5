"""B200-native Chamfer / density-aware Chamfer (DCD) hot path of U-RED.

Drop-in surface (same names and signatures as the reference):
    chamfer_3DDist()(x, y) -> (dist1, dist2, idx1, idx2)
    calc_dcd(x, gt, alpha=1000, n_lambda=1, return_raw=False, non_reg=False)
    calc_cd(output, gt, calc_f1=False, return_raw=False, normalize=False, separate=False)
    cd, fscore                      (Density_aware_Chamfer_Distance/utils_v2/metrics/__init__.py)
    ChamferLoss, chamfer_distance2, compute_cm_loss      (loss/chamfer_loss.py)
plus the batched retrieval API in .retrieval.  Everything runs on the hand-written sm_100a
kernels behind include/ured_chamfer.h; there is no CPU, PyTorch or Triton fallback.
"""
from . import _native
from ._native import NativeLibraryError, build_native
from .dist_chamfer_3D import chamfer_3DDist, chamfer_3DFunction, nn_forward, nn_backward, check_finite
from .dist_chamfer_3D import chamfer_3DDist as cd
from .model_utils import calc_cd, calc_dcd, chamfer_ragged, fscore, fscore_fused, torch_epilogue
from .chamfer_loss import ChamferLoss, chamfer_distance2, compute_cm_loss
from . import retrieval
from . import compat
from .graphed import GraphedDCD
from .knn import knn1_points, residual_retrieval_loss
from .retrieval import (RetrievalEngine, PendingResult, PackedClouds, score_all_pairs, write_pair_pickles, score_candidates, score_library, topk_smallest, retrieve,
                        retrieve_sharded, shard_bounds, merge_topk, gather_and_merge)
from .exchange import PeerExchange
from .emd_module import emdFunction, emdModule, calc_emd, rerank_emd

__all__ = [
    "chamfer_3DDist", "chamfer_3DFunction", "nn_forward", "nn_backward", "check_finite", "cd", "fscore", "calc_cd", "calc_dcd", "chamfer_ragged",
    "ChamferLoss", "chamfer_distance2", "compute_cm_loss",
    "knn1_points", "residual_retrieval_loss", "PackedClouds", "RetrievalEngine", "PendingResult", "GraphedDCD", "score_candidates", "score_library", "score_all_pairs", "write_pair_pickles", "topk_smallest", "retrieve",
    "retrieve_sharded", "shard_bounds", "merge_topk", "gather_and_merge", "PeerExchange", "fscore_fused", "torch_epilogue", "emdFunction", "emdModule", "calc_emd", "rerank_emd",
    "NativeLibraryError", "build_native",
]
