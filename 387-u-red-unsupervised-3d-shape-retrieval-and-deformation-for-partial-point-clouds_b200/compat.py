"""Module aliases that let the reference's own files import the B200 path without being edited.

    import ured_b200
    ured_b200.compat.install()
    # from here on, inside the reference tree:
    #   from Density_aware_Chamfer_Distance.utils_v2.metrics import cd, fscore          (metrics/__init__.py:1-6)
    #   from Density_aware_Chamfer_Distance.utils_v2.model_utils import calc_dcd, calc_cd (engine/geometry_utils.py:9)
    #   from Shape_Measure.distance import ChamferLoss                                   (loss/chamfer_loss.py:3)
    #   from pytorch3d.loss import chamfer_distance ; from pytorch3d.ops import knn_points (loss/chamfer_loss.py:1, loss/basic_loss.py)
    # resolve to this package.

`emd` / `calc_emd` resolve to the B200 auction-EMD kernel (emd_module.py).  `Shape_Measure.distance.EMDLoss` -- a module
that is neither vendored nor pinned in the reference, with an unknown call contract -- stays a placeholder that raises
when CALLED, so importing modules that merely mention it keeps working.

`Density_aware_Chamfer_Distance.*` and `Shape_Measure.*` are ALWAYS aliased: they are exactly what this package
replaces, and with the reference root on sys.path (the documented flow) the genuine DCD package is importable -- leaving
it in place would silently JIT-build and run the reference's CUDA op.  The alias packages keep the genuine package's
search path, so its other submodules (models, cfgs, ...) still import from disk.  Only a genuinely installed
`pytorch3d` is left alone unless ``force=True`` (a warning says so).
"""
import warnings
import importlib.util
import sys
import types

import torch

from . import knn as _knn
from . import model_utils as _mu
from .chamfer_loss import ChamferLoss
from .dist_chamfer_3D import chamfer_3DDist, chamfer_3DFunction


class _OutOfScope:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise NotImplementedError(f"{self._name} is outside the B200 Chamfer/DCD hot path (EMD auction op, SURVEY.md section 2)")


def chamfer_distance(x, y, batch_reduction="mean", point_reduction="mean"):
    """The subset of pytorch3d.loss.chamfer_distance the reference uses (loss/chamfer_loss.py:1,
    engine/geometry_utils.py:65-67): squared-L2 Chamfer of dense clouds, no normals/lengths/weights.
    Returns (loss, None) like pytorch3d's (loss, loss_normals)."""
    d1, d2, _, _ = chamfer_3DDist()(x.float(), y.float())
    if point_reduction == "mean":
        cham = d1.mean(1) + d2.mean(1)
    elif point_reduction == "sum":
        cham = d1.sum(1) + d2.sum(1)
    else:
        raise ValueError("point_reduction must be 'mean' or 'sum'")
    if batch_reduction == "mean":
        cham = cham.mean()
    elif batch_reduction == "sum":
        cham = cham.sum()
    elif batch_reduction is not None:
        raise ValueError("batch_reduction must be 'mean', 'sum' or None")
    return cham, None


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False, **_unused):
    """pytorch3d.ops.knn_points for K=1 (the only use on the path: loss/basic_loss.py:257)."""
    if K != 1:
        raise NotImplementedError("only K=1 is provided on the B200 path")
    if lengths1 is not None:
        raise NotImplementedError("lengths1 is not supported (the reference never passes it)")
    dists, idx, nn = _knn.knn1_points(p1, p2, lengths2=lengths2, return_nn=return_nn)
    import collections
    return collections.namedtuple("KNN", "dists idx knn")(dists, idx, nn)


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__ured_b200_compat__ = True
    return mod


def _really_installed(name):
    mod = sys.modules.get(name)
    if mod is not None:
        return not getattr(mod, "__ured_b200_compat__", False)
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install(force=False):
    """Register the alias modules in sys.modules; returns the list of names installed."""
    from . import emd_module as _emd
    emd = _emd.emdModule
    mods = {
        "Density_aware_Chamfer_Distance": {},
        "Density_aware_Chamfer_Distance.utils_v2": {},
        "Density_aware_Chamfer_Distance.utils_v2.metrics": dict(cd=chamfer_3DDist, fscore=_mu.fscore, emd=emd, __all__=["cd", "fscore", "emd"]),
        "Density_aware_Chamfer_Distance.utils_v2.metrics.CD": dict(cd=chamfer_3DDist, fscore=_mu.fscore),
        "Density_aware_Chamfer_Distance.utils_v2.metrics.EMD": dict(emd=_emd.emdModule),
        "Density_aware_Chamfer_Distance.utils_v2.metrics.EMD.emd_module": dict(emdFunction=_emd.emdFunction, emdModule=_emd.emdModule),
        "Density_aware_Chamfer_Distance.utils_v2.metrics.CD.chamfer3D": {},
        "Density_aware_Chamfer_Distance.utils_v2.metrics.CD.chamfer3D.dist_chamfer_3D": dict(chamfer_3DDist=chamfer_3DDist, chamfer_3DFunction=chamfer_3DFunction),
        "Density_aware_Chamfer_Distance.utils_v2.model_utils": dict(calc_dcd=_mu.calc_dcd, calc_cd=_mu.calc_cd, calc_emd=_emd.calc_emd,
                                                                   cd=chamfer_3DDist, fscore=_mu.fscore, emd=emd),
        "Shape_Measure": {},
        "Shape_Measure.distance": dict(ChamferLoss=ChamferLoss, EMDLoss=_OutOfScope("EMDLoss")),
        "pytorch3d": {},
        "pytorch3d.loss": dict(chamfer_distance=chamfer_distance),
        "pytorch3d.ops": dict(knn_points=knn_points),
    }
    done = []
    keep_genuine = set()
    if not force and _really_installed("pytorch3d"):
        keep_genuine.add("pytorch3d")
        warnings.warn("ured_b200.compat.install(): a genuine pytorch3d is installed and is left in place; pass force=True to route "
                      "pytorch3d.loss.chamfer_distance / pytorch3d.ops.knn_points to the B200 path", stacklevel=2)
    for name, attrs in mods.items():
        top = name.split(".")[0]
        if top in keep_genuine:
            continue
        mod = _module(name, **attrs)
        if attrs == {}:
            # package: keep the genuine package's search path (if there is one) for the submodules we do not alias
            path = []
            if not getattr(sys.modules.get(name), "__ured_b200_compat__", False):
                try:
                    spec = importlib.util.find_spec(name) if name not in sys.modules else getattr(sys.modules[name], "__spec__", None)
                    if spec is not None and spec.submodule_search_locations:
                        path = list(spec.submodule_search_locations)
                except (ImportError, ValueError, AttributeError):
                    path = []
            else:
                path = list(getattr(sys.modules[name], "__path__", []))
            mod.__path__ = path
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
        done.append(name)
    return done
