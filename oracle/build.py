"""Build recipes for the oracle (test infrastructure only).

build_oracle()  gcc  oracle/chamfer_oracle.c            -> oracle/liboracle_chamfer.so
build_ref()     nvcc the UNMODIFIED reference op, read in place from /root/reference
                (chamfer3D.cu + chamfer_cuda.cpp, as dist_chamfer_3D.py:11-16 JIT-loads them)
                                                        -> oracle/_ref/chamfer_3D_ref*.so
Only outputs are written under oracle/_ref/ (git-ignored, shipped to the GPU box); no reference
source is copied into this repository.  build_ref() is a no-op when /root/reference is absent
(on the GPU box the prebuilt module is used).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle_chamfer.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/Density_aware_Chamfer_Distance/utils_v2/metrics/CD/chamfer3D"
REF_NAME = "chamfer_3D_ref"


def build_oracle(force=False):
    src = os.path.join(HERE, "chamfer_oracle.c")
    if not force and os.path.exists(ORACLE_SO) and os.path.getmtime(ORACLE_SO) >= os.path.getmtime(src):
        return ORACLE_SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", ORACLE_SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return ORACLE_SO


EMD_SO = os.path.join(HERE, "liboracle_emd.so")


def build_emd_oracle(force=False):
    """gcc oracle/emd_oracle.c -> oracle/liboracle_emd.so (CPU restatement of the reference's auction EMD)."""
    src = os.path.join(HERE, "emd_oracle.c")
    if not force and os.path.exists(EMD_SO) and os.path.getmtime(EMD_SO) >= os.path.getmtime(src):
        return EMD_SO
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", EMD_SO, src, "-lm"], check=True)
    return EMD_SO


SCREEN_SO = os.path.join(HERE, "libscreen_model.so")


def build_screen_model(force=False):
    """gcc oracle/screen_model.c -> oracle/libscreen_model.so (CPU emulation of the screening invariant)."""
    src = os.path.join(HERE, "screen_model.c")
    if not force and os.path.exists(SCREEN_SO) and os.path.getmtime(SCREEN_SO) >= os.path.getmtime(src):
        return SCREEN_SO
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", SCREEN_SO, src, "-lm"], check=True)
    return SCREEN_SO


def ref_module_path():
    if not os.path.isdir(REF_DIR):
        return None
    for f in sorted(os.listdir(REF_DIR)):
        if f.startswith(REF_NAME) and f.endswith(".so"):
            return os.path.join(REF_DIR, f)
    return None


def build_ref(force=False):
    """Compile the reference CUDA op for sm_100a from its sources where they lie."""
    if not os.path.isdir(REF_SRC):
        return ref_module_path()
    if not force and ref_module_path():
        return ref_module_path()
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name=REF_NAME,
         sources=[os.path.join(REF_SRC, "chamfer_cuda.cpp"), os.path.join(REF_SRC, "chamfer3D.cu")],
         build_directory=REF_DIR, verbose=False, is_python_module=True)
    return ref_module_path()


REF_EMD_SRC = "/root/reference/Density_aware_Chamfer_Distance/utils_v2/metrics/EMD"
REF_EMD_NAME = "emd_ref"
REF_EMD_DIR = os.path.join(REF_DIR, "emd")


def ref_emd_module_path():
    if not os.path.isdir(REF_EMD_DIR):
        return None
    for f in sorted(os.listdir(REF_EMD_DIR)):
        if f.startswith(REF_EMD_NAME) and f.endswith(".so"):
            return os.path.join(REF_EMD_DIR, f)
    return None


def build_ref_emd(force=False):
    """Compile the reference's auction-EMD op (emd.cpp + emd_cuda.cu, as EMD/emd_module.py:31-36 JIT-loads them) for
    sm_100a from its sources where they lie -> oracle/_ref/emd/emd_ref.so (the checker of the EMD re-rank)."""
    if not os.path.isdir(REF_EMD_SRC):
        return ref_emd_module_path()
    if not force and ref_emd_module_path():
        return ref_emd_module_path()
    os.makedirs(REF_EMD_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name=REF_EMD_NAME,
         sources=[os.path.join(REF_EMD_SRC, "emd.cpp"), os.path.join(REF_EMD_SRC, "emd_cuda.cu")],
         build_directory=REF_EMD_DIR, verbose=False, is_python_module=True)
    return ref_emd_module_path()


def load_ref_emd():
    path = ref_emd_module_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(REF_EMD_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref():
    """Import the prebuilt reference op (needs torch; callable only with a GPU)."""
    path = ref_module_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(REF_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
