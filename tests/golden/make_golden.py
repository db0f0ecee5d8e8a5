"""Generate the committed golden vectors from the REFERENCE's own Python code.

Run in the authoring container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Writes
  tests/golden/chamfer_ref_python.npz   outputs of the reference's float64 python Chamfer
        (Density_aware_Chamfer_Distance/utils_v2/metrics/CD/chamfer_python.py:18-39, imported as is)
        on the reference unit test's shapes (unit_test.py:15-16) and on shape-like clouds;
  tests/golden/dcd_ref_model_utils.npz  outputs and gradients of the reference's calc_dcd / calc_cd
        (Density_aware_Chamfer_Distance/utils_v2/model_utils.py:13-70, executed UNMODIFIED) with the
        package import `...utils_v2.metrics` satisfied by a stub whose `cd` is the C oracle -- the
        real `cd` JIT-builds a CUDA extension at import and needs a GPU (metrics/__init__.py:1-2).
The second file therefore pins the torch-op body of calc_dcd/calc_cd, the first one pins the
oracle's nearest-neighbour search against an independent float64 implementation.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
CD_DIR = os.path.join(REF, "Density_aware_Chamfer_Distance/utils_v2/metrics/CD")


def load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def clouds(seed, b, n, kind):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "U":  # the reference test's distribution, unit_test.py:15-16
        return torch.rand(b, n, 3, generator=g)
    x = torch.randn(b, n, 3, generator=g)  # "S": centred, scaled into the unit ball (geometry_utils.py:88-94)
    x = x - x.mean(1, keepdim=True)
    return x / x.norm(dim=2).amax(1).view(b, 1, 1)


def main():
    torch.set_num_threads(4)
    ref_py = load_file("ref_chamfer_python", os.path.join(CD_DIR, "chamfer_python.py"))
    ref_fscore = load_file("ref_fscore", os.path.join(CD_DIR, "fscore.py"))

    out = {}
    cases = [("unit_test", "U", 4, 100, 200, 0), ("timing", "U", 2, 2000, 1000, 10), ("shape", "S", 3, 777, 1301, 20),
             ("chair", "S", 2, 2048, 2048, 30)]
    for name, kind, b, n, m, seed in cases:
        p1, p2 = clouds(seed, b, n, kind), clouds(seed + 1, b, m, kind)
        d1, d2, i1, i2 = ref_py.distChamfer(p1, p2)
        out[f"{name}_xyz1"], out[f"{name}_xyz2"] = p1.numpy(), p2.numpy()
        out[f"{name}_dist1"], out[f"{name}_dist2"] = d1.numpy(), d2.numpy()
        out[f"{name}_idx1"], out[f"{name}_idx2"] = i1.numpy(), i2.numpy()
    np.savez_compressed(os.path.join(HERE, "chamfer_ref_python.npz"), **out)

    # ---- reference calc_dcd / calc_cd, unmodified, over the oracle Chamfer ------------------
    from oracle import torch_path

    class OracleCD(torch.nn.Module):
        def forward(self, a, b):
            return torch_path.oracle_cd(a, b)

    pkg = "Density_aware_Chamfer_Distance.utils_v2.metrics"
    for i, part in enumerate(pkg.split(".")):
        sys.modules.setdefault(".".join(pkg.split(".")[: i + 1]), types.ModuleType(part))
    stub = sys.modules[pkg]
    stub.cd, stub.emd, stub.fscore = OracleCD, None, ref_fscore.fscore
    ref_mu = load_file("ref_model_utils", os.path.join(REF, "Density_aware_Chamfer_Distance/utils_v2/model_utils.py"))

    out = {}
    dcd_cases = [("default", 1000, 1, False, 3, 512, 512, 40), ("pcn", 200, 0.5, False, 2, 700, 1024, 50),
                 ("vrc_nonreg", 40, 0.5, True, 2, 1024, 300, 60), ("lambda2", 50, 2, False, 2, 256, 384, 70)]
    for name, alpha, lam, non_reg, b, n_x, n_gt, seed in dcd_cases:
        x = clouds(seed, b, n_x, "S").requires_grad_()
        noise = 0.02 * torch.randn(b, n_gt, 3, generator=torch.Generator().manual_seed(seed + 5))
        gt = (clouds(seed + 1, b, n_gt, "S") * 0.9 + noise).requires_grad_()
        loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = ref_mu.calc_dcd(x, gt, alpha=alpha, n_lambda=lam,
                                                                  return_raw=True, non_reg=non_reg)
        w = torch.linspace(0.5, 1.5, b)
        (loss * w).sum().backward()
        g_x_loss, g_gt_loss = x.grad.clone(), gt.grad.clone()
        x.grad = None; gt.grad = None
        cd_p2, cd_t2, f1 = ref_mu.calc_cd(x, gt, calc_f1=True)
        ((cd_p2 + 3 * cd_t2) * w).sum().backward()
        out[f"{name}_meta"] = np.array([alpha, lam, float(non_reg)], np.float64)
        for k, v in dict(x=x, gt=gt, loss=loss, cd_p=cd_p, cd_t=cd_t, dist1=dist1, dist2=dist2, idx1=idx1, idx2=idx2,
                         g_x_loss=g_x_loss, g_gt_loss=g_gt_loss, g_x_cd=x.grad, g_gt_cd=gt.grad, f1=f1, w=w).items():
            out[f"{name}_{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "dcd_ref_model_utils.npz"), **out)
    print("golden vectors written")


if __name__ == "__main__":
    main()
