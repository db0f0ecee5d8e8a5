"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and validates arguments
before touching the device (no compute calls here -- there is no GPU in this suite)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(ured_\w+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported(ured):
    lib = ctypes.CDLL(ured._native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert set(syms) == set(ured._native.EXPORTED_SYMBOLS), "ctypes binding and header disagree"


def test_abi_version_and_sizes(ured):
    lib = ured._native.load()
    assert lib.ured_abi_version() == 1
    # per cloud: 4 arrays of np floats + a 32-float tail holding the max norm; 256-aligned in total
    assert lib.ured_packed_bytes(2, 2048) == 2 * (4 * 2048 + 32) * 4
    assert lib.ured_packed_bytes(1, 100) == (4 * 128 + 32) * 4 + 128  # padded to 32 points, rounded up to 256 bytes
    assert lib.ured_chamfer_workspace_bytes(3, 100, 200) == lib.ured_packed_bytes(3, 100) + lib.ured_packed_bytes(3, 200) + lib.ured_nn_scratch_bytes(3, 100, 200)
    assert lib.ured_nn_scratch_bytes(5000, 2048, 2048) == 0           # many waves of CTAs (in either call flavour): no candidate splitting
    assert lib.ured_nn_scratch_bytes(16, 16384, 16384) == 8 * 16 * 32768 * 8  # dense clouds: 8 splits (2048-candidate ranges) of partial (d, idx)


def test_argument_errors_do_not_need_a_device(ured):
    lib = ured._native.load()
    E_NULL, E_SHAPE, E_RANGE = -1, -2, -4
    assert lib.ured_chamfer_forward(None, None, 2, 8, 8, None, None, None, None, None, None, None, 0, 0, None) == E_NULL
    assert b"NULL" in lib.ured_last_error_string()
    assert lib.ured_chamfer_forward(None, None, -1, 8, 8, None, None, None, None, None, None, None, 0, 0, None) == E_SHAPE
    assert lib.ured_nn_packed(None, None, 8, None, None, 8, 4, 0, 4, None, None, None, None, None, None, None, 0, 0, None) == E_SHAPE
    assert lib.ured_topk_smallest(None, 1, 5, 6, 0, None, None, None) == E_RANGE
    assert lib.ured_dcd_forward(None, None, None, None, 1, 8, 8, 1, 1, None, None, 1.0, 1.0, 1.0, 1.0, 0, None, None, None, None, None, None) == E_NULL
    # empty batches are a successful no-op
    assert lib.ured_chamfer_forward(None, None, 0, 8, 8, None, None, None, None, None, None, None, 0, 0, None) == 0
    assert lib.ured_topk_smallest(None, 0, 5, 2, 0, None, None, None) == 0
    # round-2 entry points: the peer exchange and the auction EMD
    assert lib.ured_xchg_bytes(8, 64, 10) > 0 and lib.ured_xchg_bytes(8, 64, 10) % 256 == 0
    assert lib.ured_xchg_bytes(17, 1, 10) == 0 and lib.ured_xchg_bytes(8, 513, 10) == 0 and lib.ured_xchg_bytes(8, 1, 65) == 0 and lib.ured_xchg_bytes(0, 1, 1) == 0
    assert lib.ured_topk_exchange(None, 4, 20, 10, 0, None, 2, 0, 4, None, None, 100, None) == E_NULL
    assert lib.ured_topk_exchange(None, 4, 20, 10, 0, None, 17, 0, 4, None, None, 100, None) == E_RANGE     # world <= 16
    assert lib.ured_topk_exchange(None, 4, 20, 65, 0, None, 2, 0, 4, None, None, 100, None) == E_RANGE      # k <= 64
    assert lib.ured_topk_exchange(None, 0, 20, 10, 0, None, 2, 0, 0, None, None, 100, None) == 0           # no query rows: nothing to do
    assert lib.ured_emd_workspace_bytes(20, 2048) >= 20 * 2048 * 7 * 4 and lib.ured_emd_workspace_bytes(20, 2048) % 256 == 0
    assert lib.ured_emd_forward(None, None, 2, 64, 0.005, 10, None, None, None, 0, None) == E_NULL
    assert lib.ured_emd_forward(None, None, -2, 64, 0.005, 10, None, None, None, 0, None) == E_SHAPE
    assert lib.ured_emd_forward(None, None, 0, 64, 0.005, 10, None, None, None, 0, None) == 0
    assert lib.ured_emd_backward(None, None, 2, 64, None, None, None, None) == E_NULL
    assert lib.ured_nn_backward_one_direction(None, None, 2, 8, 8, None, None, None, None, None, None) == E_NULL
    assert lib.ured_probe_ffma(None, 1, 1, None, None) == E_SHAPE


def test_no_cpu_fallback(ured):
    import torch
    x = torch.rand(1, 8, 3)
    with pytest.raises(RuntimeError, match="GPU tensors only"):
        ured.chamfer_3DDist()(x, x)
    with pytest.raises(RuntimeError, match="GPU tensors only"):
        ured.calc_dcd(x, x)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no file of the shipped package may mention it."""
    pkg_dir = os.path.join(ROOT, "387-u-red-unsupervised-3d-shape-retrieval-and-deformation-for-partial-point-clouds_b200")
    for path in glob.glob(os.path.join(pkg_dir, "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".h", ".cpp")):
            text = open(path).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "liboracle" not in text, path


def test_header_is_plain_c(tmp_path):
    """include/ured_chamfer.h must be consumable by a C compiler (cgo/JNI/ctypes-style FFI users), not only by nvcc."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "ured_chamfer.h"\n'
                   'int probe(void) { return ured_abi_version() == URED_ABI_VERSION && URED_E_NULL < 0 && (URED_FLAG_EXACT_ONLY | URED_FLAG_NON_REG | URED_FLAG_ONE_DIRECTION) == 7u; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "use_header.o")],
                   check=True)


def test_launch_shape_heuristic_invariants(ured):
    """ured_nn_launch_shape / ured_nn_scratch_bytes expose the launch plan: 1, 2, 4 or 8 candidate splits, never below 512
    candidates per split, clouds of >= 4096 candidates cut into ranges of >= 2048, only the last partial wave split for
    mid-size launches, nothing split once the grid is many waves long, scratch consistent with the plan."""
    import ctypes
    lib = ured._native.load()
    FP32 = ured._native.URED_FLAG_FP32_SCREEN
    for B in [1, 2, 7, 16, 32, 100, 125, 640, 5000]:
        for n1, n2 in [(1, 1), (100, 200), (511, 4096), (1024, 1024), (2048, 2048), (2000, 1000), (16384, 16384), (4096, 100000)]:
            v, q, t, ns, items, split = (ctypes.c_int() for _ in range(6))
            # the tensor-core plan (default): whole groups of 128-query tiles per item, candidate ranges of 2048, no scratch use
            assert lib.ured_nn_launch_shape(B, n1, n2, 0, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns),
                                            ctypes.byref(items), ctypes.byref(split)) == 0
            assert v.value == 100 and q.value % 128 == 0 and 128 <= q.value <= 2048 and split.value == 0
            assert ns.value == -(-max(n1, n2) // 2048) and items.value == B * (-(-n1 // q.value) + -(-n2 // q.value))
            # the FP32-pipe plan
            assert lib.ured_nn_launch_shape(B, n1, n2, FP32, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns),
                                            ctypes.byref(items), ctypes.byref(split)) == 0
            nsplit = ns.value
            sb = lib.ured_nn_scratch_bytes(B, n1, n2)
            assert nsplit in (1, 2, 4, 8) and q.value % t.value == 0 and q.value in (256, 512, 1024)
            assert items.value == B * (-(-n1 // q.value) + -(-n2 // q.value)) and 0 <= split.value <= items.value
            assert (split.value == 0) == (nsplit == 1)
            assert sb % 256 == 0 and sb >= (nsplit * split.value * q.value * 8 if split.value else 0)
            if nsplit > 1:
                assert min(n1, n2) // nsplit >= 512                       # a split never gets fewer than 512 candidates
            if min(n1, n2) < 4096 and items.value >= 148 * 5 * 4:
                assert split.value == 0                                   # many waves of CTAs: no splitting at all
            if min(n1, n2) < 4096 and 148 * 2 <= items.value and split.value:
                assert split.value < 148 * 5 and nsplit <= 4              # mid-size launch: only the last partial wave is cut
            if min(n1, n2) >= 4096 and items.value * nsplit >= 148 * 2:
                assert split.value == items.value and (min(n1, n2) // nsplit >= 2048 or nsplit == 8)   # big clouds: 2048-candidate ranges
            assert lib.ured_chamfer_workspace_bytes(B, n1, n2) == lib.ured_packed_bytes(B, n1) + lib.ured_packed_bytes(B, n2) + sb
    # the tail rule (split only the last partial wave) is an experiment knob, off by default: the cfg3 shard of 125 shapes
    # (1000 work items on 740 CTA slots) launches unsplit
    v, q, t, ns, items, split = (ctypes.c_int() for _ in range(6))
    lib.ured_nn_launch_shape(125, 2048, 2048, FP32, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns), ctypes.byref(items), ctypes.byref(split))
    assert (items.value, split.value, ns.value) == (1000, 0, 1)
    assert lib.ured_nn_scratch_bytes(0, 8, 8) == 0 and lib.ured_nn_scratch_bytes(4, 0, 8) == 0


def test_tensor_core_launch_plan_for_the_baseline_shapes(ured):
    """The work-item rule of the tensor-core kernel: whole clouds per item when that keeps the 148 persistent CTAs
    busy, smaller query groups for small batches, candidate ranges of 2048 for large clouds."""
    import ctypes
    lib = ured._native.load()

    def plan(B, n1, n2, flags=0):
        v, q, t, ns, items, split = (ctypes.c_int() for _ in range(6))
        assert lib.ured_nn_launch_shape(B, n1, n2, flags, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns),
                                        ctypes.byref(items), ctypes.byref(split)) == 0
        return v.value, q.value, t.value, ns.value, items.value, split.value

    assert plan(32, 2048, 2048) == (100, 1024, 448, 1, 128, 0)        # cfg1: 8 query tiles per item, one round of 128 items
    assert plan(640, 2048, 2048) == (100, 2048, 448, 1, 1280, 0)      # cfg2: a whole cloud per item
    assert plan(16, 16384, 16384) == (100, 2048, 448, 8, 256, 0)      # cfg4: 16 tiles per item, eight candidate ranges each
    assert plan(125, 2048, 2048)[4] == 250                            # one cfg3 shard of an 8-GPU run
    assert plan(1, 2048, 2048)[1:5] == (128, 448, 1, 32)              # one pair: every query tile its own item
    one_dir = plan(3, 700, 3072, ured._native.URED_FLAG_ONE_DIRECTION)
    assert one_dir[0] == 100 and one_dir[3] == 2 and one_dir[4] == 3 * -(-(-(-700 // 128)) // (one_dir[1] // 128))
    assert plan(640, 2048, 2048, ured._native.URED_FLAG_FP32_SCREEN)[0] == 0 and plan(640, 2048, 2048, ured._native.URED_FLAG_EXACT_ONLY)[0] == 0
