"""Importable alias for the package directory (its mandated name is not a Python identifier).

    import ured_b200
    ured_b200.calc_dcd(x, gt)
"""
import importlib
import os
import sys

PACKAGE_DIR_NAME = "387-u-red-unsupervised-3d-shape-retrieval-and-deformation-for-partial-point-clouds_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(PACKAGE_DIR_NAME)
sys.modules[__name__] = _pkg
