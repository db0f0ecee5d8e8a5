"""CPU: the alias modules let the REFERENCE's own caller files import unmodified (only where /root/reference exists)."""
import importlib.util
import os
import sys

import pytest

REF = "/root/reference"


@pytest.fixture()
def installed(ured):
    before = dict(sys.modules)
    names = ured.compat.install()
    yield ured, names
    for n in list(sys.modules):
        if n not in before:
            del sys.modules[n]


def test_aliases_resolve_to_this_package(installed):
    ured, names = installed
    assert "Shape_Measure.distance" in names and "Density_aware_Chamfer_Distance.utils_v2.model_utils" in names
    from Density_aware_Chamfer_Distance.utils_v2.metrics import cd, fscore
    from Density_aware_Chamfer_Distance.utils_v2.model_utils import calc_cd, calc_dcd, calc_emd
    from Shape_Measure.distance import ChamferLoss, EMDLoss
    assert cd is ured.chamfer_3DDist and calc_dcd is ured.calc_dcd and calc_cd is ured.calc_cd and fscore is ured.fscore
    assert ChamferLoss is ured.ChamferLoss
    assert calc_emd is ured.calc_emd                              # the auction EMD is part of the B200 path now
    from Density_aware_Chamfer_Distance.utils_v2.metrics import emd
    assert emd is ured.emdModule
    with pytest.raises(NotImplementedError):                      # Shape_Measure.EMDLoss: absent from the reference, contract unknown
        EMDLoss()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the authoring container")
def test_reference_caller_files_import_unmodified(installed):
    """loss/chamfer_loss.py of the reference (imports pytorch3d + Shape_Measure, both absent here) loads over the aliases."""
    ured, _ = installed
    spec = importlib.util.spec_from_file_location("ref_chamfer_loss", os.path.join(REF, "loss/chamfer_loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ChamferLoss is ured.ChamferLoss
    assert callable(mod.chamfer_distance2) and callable(mod.compute_cm_loss) and mod.chamfer_distance is ured.compat.chamfer_distance
    import torch
    with pytest.raises(RuntimeError, match="GPU tensors only"):   # the reference's code reaches our op (no CPU fallback)
        mod.chamfer_distance2(torch.rand(1, 8, 3), torch.rand(1, 8, 3))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the authoring container")
def test_aliases_win_with_the_reference_root_on_sys_path(ured):
    """INTEGRATION.md's flow: the reference root is on sys.path, so the genuine DCD package is importable -- the aliases
    must still take over (otherwise engine/geometry_utils.py would JIT-build and run the reference's CUDA op)."""
    before = dict(sys.modules)
    sys.path.insert(0, REF)
    try:
        names = ured.compat.install()
        assert "Density_aware_Chamfer_Distance.utils_v2.model_utils" in names and "Shape_Measure.distance" in names
        from Density_aware_Chamfer_Distance.utils_v2.model_utils import calc_dcd
        from Density_aware_Chamfer_Distance.utils_v2.metrics import cd
        import Density_aware_Chamfer_Distance
        assert calc_dcd is ured.calc_dcd and cd is ured.chamfer_3DDist
        # the alias package still finds the genuine package's other submodules on disk
        assert any(p.startswith(REF) for p in Density_aware_Chamfer_Distance.__path__)
        assert importlib.util.find_spec("Density_aware_Chamfer_Distance.models") is not None
    finally:
        sys.path.remove(REF)
        for n in list(sys.modules):
            if n not in before:
                del sys.modules[n]
