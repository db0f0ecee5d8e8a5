"""The C ABI used from plain C (examples/c_abi_demo.c): links without Python/torch; on a GPU it must reproduce the
reference arithmetic bit for bit."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_demo(ured, out_dir):
    pkg = os.path.dirname(ured._native.LIB_PATH)
    exe = os.path.join(out_dir, "c_abi_demo")
    subprocess.run(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(CUDA, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
                    "-L", pkg, "-lured_chamfer", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-o", exe], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.pathsep.join([pkg, os.path.join(CUDA, "lib64"), os.environ.get("LD_LIBRARY_PATH", "")]))
    return exe, env


@pytest.mark.skipif(shutil.which("gcc") is None or not os.path.isdir(os.path.join(CUDA, "include")), reason="needs gcc and the CUDA headers")
def test_demo_links_from_plain_c(ured, tmp_path):
    exe, _ = build_demo(ured, str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_demo_runs_bit_exact(ured, tmp_path):
    exe, env = build_demo(ured, str(tmp_path))
    out = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches = 0" in out.stdout and "bad-call code = -1" in out.stdout
