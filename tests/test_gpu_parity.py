"""GPU: parity of the CUDA path (through the C ABI) with the oracle.

Bars (BASELINE.json north_star): idx1/idx2 and rankings bit-exact; distances bit-exact (they are
the same float32 difference-form values); gradients and DCD values within 1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5  # north_star tolerance for floating-point outputs


def dev(t):
    return t.cuda()


KERNELS = [False, "fp32", True]          # screening on the tensor cores (default) / on the FP32 pipes / difference form on every pair
KERNEL_IDS = ["tensor", "fp32", "exact"]


def run_fwd(ured, a, b, exact_only):
    out = ured.nn_forward(dev(a).contiguous(), dev(b).contiguous(), exact_only=exact_only is True, fp32_screen=exact_only == "fp32")
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def assert_bit_exact(got, want, tag=""):
    for g, w, name in zip(got, want, ["dist1", "dist2", "idx1", "idx2"]):
        if g.dtype == np.float32:
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), f"{tag} {name} differs: max abs {np.abs(g - w).max()}"
        else:
            assert np.array_equal(g, w), f"{tag} {name}: {np.sum(g != w)} indices differ"


SHAPES = [  # (B, N, M, kind)
    (4, 100, 200, "U"),     # the reference unit test, unit_test.py:15-16
    (2, 2000, 1000, "U"),   # the reference timing shapes, unit_test.py:39-40
    (1, 1, 1, "U"), (1, 1, 37, "U"), (3, 37, 1, "U"),
    (2, 31, 33, "U"), (2, 513, 1025, "S"), (1, 1024, 1023, "S"),
    (5, 2048, 2048, "S"),   # chair
    (1, 5000, 3000, "U"),   # several 1024-candidate TMA tiles, ragged tail
    (2, 1024, 2048, "S"),   # per-part loss shape, loss/chamfer_loss.py:23
]


@pytest.mark.parametrize("exact_only", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("B,N,M,kind", SHAPES)
def test_forward_bit_exact_vs_oracle(ured, oracle, B, N, M, kind, exact_only):
    a, b = make_clouds(0, B, N, kind), make_clouds(1, B, M, kind)
    want = oracle.c.chamfer_forward(a.numpy(), b.numpy())
    assert_bit_exact(run_fwd(ured, a, b, exact_only), want, f"{B}x{N}x{M}")


@pytest.mark.parametrize("exact_only", KERNELS, ids=KERNEL_IDS)
def test_forward_adversarial_ties_and_duplicates(ured, oracle, exact_only):
    g = torch.Generator().manual_seed(7)
    lattice_a = torch.randint(0, 4, (2, 700, 3), generator=g).float()          # masses of exact ties
    lattice_b = torch.randint(0, 4, (2, 900, 3), generator=g).float()
    assert_bit_exact(run_fwd(ured, lattice_a, lattice_b, exact_only), oracle.c.chamfer_forward(lattice_a.numpy(), lattice_b.numpy()), "lattice")
    x = make_clouds(3, 2, 600, "S")
    dup = torch.cat([x, x[:, :300]], 1)                                          # duplicated candidates
    assert_bit_exact(run_fwd(ured, x, dup, exact_only), oracle.c.chamfer_forward(x.numpy(), dup.numpy()), "dup")
    same = make_clouds(4, 1, 1500, "U")
    got = run_fwd(ured, same, same, exact_only)                                  # identical clouds: d = 0, idx = self
    assert_bit_exact(got, oracle.c.chamfer_forward(same.numpy(), same.numpy()), "self")
    assert (got[0] == 0).all() and np.array_equal(got[2][0], np.arange(1500))
    far = make_clouds(5, 1, 800, "S") + 1000.0                                   # large offset: screening bound must widen
    far_b = make_clouds(6, 1, 800, "S") + 1000.0
    assert_bit_exact(run_fwd(ured, far, far_b, exact_only), oracle.c.chamfer_forward(far.numpy(), far_b.numpy()), "offset")
    tiny = make_clouds(8, 1, 500, "S") * 1e-20                                   # products underflow to denormals / zero
    tiny_b = make_clouds(9, 1, 400, "S") * 1e-20
    assert_bit_exact(run_fwd(ured, tiny, tiny_b, exact_only), oracle.c.chamfer_forward(tiny.numpy(), tiny_b.numpy()), "tiny")
    huge = make_clouds(10, 1, 300, "S") * 1e19                                   # squared distances overflow to +inf
    huge_b = make_clouds(11, 1, 300, "S") * 1e19
    assert_bit_exact(run_fwd(ured, huge, huge_b, exact_only), oracle.c.chamfer_forward(huge.numpy(), huge_b.numpy()), "huge")


def test_screen_and_exact_kernels_agree_at_full_size(ured):
    """cfg1-sized (B=32, 2048^2), a dense 16384^2 pair (eight candidate ranges per cloud in the tensor-core kernel) and unequal
    clouds: the three kernels must give identical bits."""
    for (B, N, M, kind) in [(32, 2048, 2048, "S"), (1, 16384, 16384, "S"), (1, 16384, 16384, "U"), (3, 4100, 2049, "S"), (150, 700, 300, "U")]:
        a, b = make_clouds(20, B, N, kind), make_clouds(21, B, M, kind)
        want = run_fwd(ured, a, b, True)
        assert_bit_exact(run_fwd(ured, a, b, False), want, f"tensor {B}x{N}x{M}")
        assert_bit_exact(run_fwd(ured, a, b, "fp32"), want, f"fp32 {B}x{N}x{M}")


def test_golden_reference_python(ured):
    """Against the committed outputs of the reference's python Chamfer: its own unit-test criterion."""
    g = np.load(os.path.join(GOLD, "chamfer_ref_python.npz"))
    for case in ["unit_test", "timing", "shape", "chair"]:
        d1, d2, i1, i2 = run_fwd(ured, torch.from_numpy(g[f"{case}_xyz1"]), torch.from_numpy(g[f"{case}_xyz2"]), False)
        assert np.mean((d1 - g[f"{case}_dist1"]) ** 2) + np.mean((d2 - g[f"{case}_dist2"]) ** 2) < 1e-8
        assert np.array_equal(i1, g[f"{case}_idx1"]) and np.array_equal(i2, g[f"{case}_idx2"])


def test_golden_reference_cuda_op(ured):
    """Against the committed outputs of the unmodified reference CUDA op on a B200 (make_golden_gpu.py): bits for dist/idx."""
    g = np.load(os.path.join(GOLD, "chamfer_ref_cuda_b200.npz"))
    for case in ["unit_test", "timing", "ragged_tail", "lattice", "chair"]:
        for exact_only in KERNELS:
            got = run_fwd(ured, torch.from_numpy(g[f"{case}_xyz1"]), torch.from_numpy(g[f"{case}_xyz2"]), exact_only)
            assert_bit_exact(got, [g[f"{case}_dist1"], g[f"{case}_dist2"], g[f"{case}_idx1"], g[f"{case}_idx2"]], case)
        xa = dev(torch.from_numpy(g[f"{case}_xyz1"])).requires_grad_()
        xb = dev(torch.from_numpy(g[f"{case}_xyz2"])).requires_grad_()
        d1, d2, _, _ = ured.chamfer_3DDist()(xa, xb)
        ((d1 * dev(torch.from_numpy(g[f"{case}_w1"]))).sum() + (d2 * dev(torch.from_numpy(g[f"{case}_w2"]))).sum()).backward()
        for got, want in [(xa.grad, g[f"{case}_grad1"]), (xb.grad, g[f"{case}_grad2"])]:
            assert np.abs(got.cpu().numpy() - want).max() / (np.abs(want).max() + 1e-30) < RTOL


def rel_err(got, want):
    scale = np.abs(want).max() + 1e-30
    return np.abs(got - want).max() / scale


@pytest.mark.parametrize("B,N,M", [(4, 100, 200), (2, 2000, 1000), (3, 2048, 2048), (1, 1, 5)])
def test_backward_vs_oracle(ured, oracle, B, N, M):
    kind = "S" if min(N, M) > 1 else "U"
    a, b = make_clouds(0, B, N, kind), make_clouds(1, B, M, kind)
    g = torch.Generator().manual_seed(2)
    w1, w2 = torch.randn(B, N, generator=g), torch.randn(B, M, generator=g)
    xa, xb = dev(a).requires_grad_(), dev(b).requires_grad_()
    d1, d2, i1, i2 = ured.chamfer_3DDist()(xa, xb)
    assert i1.dtype == torch.int32 and d1.dtype == torch.float32 and not i1.requires_grad
    ((d1 * dev(w1)).sum() + (d2 * dev(w2)).sum()).backward()
    o = oracle.c.chamfer_forward(a.numpy(), b.numpy())
    r1, r2 = oracle.c.chamfer_backward_f64(a.numpy(), b.numpy(), w1.numpy(), w2.numpy(), o[2], o[3])
    assert rel_err(xa.grad.cpu().numpy(), r1) < RTOL
    assert rel_err(xb.grad.cpu().numpy(), r2) < RTOL
    # only one output used: the other upstream gradient is None
    xa.grad = None; xb.grad = None
    d1, d2, _, _ = ured.chamfer_3DDist()(xa, xb)
    d1.sum().backward()
    r1, r2 = oracle.c.chamfer_backward_f64(a.numpy(), b.numpy(), np.ones((B, N), np.float32), np.zeros((B, M), np.float32), o[2], o[3])
    assert rel_err(xa.grad.cpu().numpy(), r1) < RTOL and rel_err(xb.grad.cpu().numpy(), r2) < RTOL


DCD_CASES = [(1000, 1, False, 3, 512, 512), (200, 0.5, False, 2, 700, 1024), (40, 0.5, True, 2, 1024, 300),
             (50, 2, False, 2, 256, 384), (30, 0.7, False, 1, 300, 300), (1000, 1, False, 2, 2048, 2048)]


@pytest.mark.parametrize("alpha,lam,non_reg,B,n_x,n_gt", DCD_CASES)
def test_calc_dcd_value_and_grad_vs_oracle(ured, oracle, alpha, lam, non_reg, B, n_x, n_gt):
    x0 = make_clouds(40, B, n_x, "S")
    gt0 = make_clouds(41, B, n_gt, "S") * 0.9 + 0.02 * torch.randn(B, n_gt, 3, generator=torch.Generator().manual_seed(9))
    w = torch.linspace(0.5, 1.5, B)
    x, gt = dev(x0).requires_grad_(), dev(gt0).requires_grad_()
    res = ured.calc_dcd(x, gt, alpha=alpha, n_lambda=lam, return_raw=True, non_reg=non_reg)
    loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = res
    xo, gto = x0.clone().requires_grad_(), gt0.clone().requires_grad_()
    ol, op, ot, od1, od2, oi1, oi2 = oracle.t.calc_dcd_oracle(xo, gto, alpha=alpha, n_lambda=lam, return_raw=True, non_reg=non_reg)
    assert np.array_equal(idx1.cpu().numpy(), oi1.numpy()) and np.array_equal(idx2.cpu().numpy(), oi2.numpy())
    assert np.array_equal(dist1.detach().cpu().numpy(), od1.detach().numpy())
    for got, want, name in [(loss, ol, "loss"), (cd_p, op, "cd_p"), (cd_t, ot, "cd_t")]:
        assert got.shape == (B,) and got.dtype == torch.float32
        assert np.allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=RTOL, atol=0), name
    (loss * dev(w)).sum().backward()
    (ol * w).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), xo.grad.numpy()) < RTOL
    assert rel_err(gt.grad.cpu().numpy(), gto.grad.numpy()) < RTOL
    # all three outputs in one loss (cd_p's sqrt has finite gradient here: no zero distances)
    x.grad = None; gt.grad = None; xo.grad = None; gto.grad = None
    l2, p2, t2 = ured.calc_dcd(x, gt, alpha=alpha, n_lambda=lam, non_reg=non_reg)
    ((l2 + 0.3 * p2 + 2.0 * t2) * dev(w)).sum().backward()
    l3, p3, t3 = oracle.t.calc_dcd_oracle(xo, gto, alpha=alpha, n_lambda=lam, non_reg=non_reg)
    ((l3 + 0.3 * p3 + 2.0 * t3) * w).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), xo.grad.numpy()) < RTOL
    assert rel_err(gt.grad.cpu().numpy(), gto.grad.numpy()) < RTOL


def test_golden_reference_model_utils(ured):
    """Against outputs of the reference's UNMODIFIED calc_dcd/calc_cd (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "dcd_ref_model_utils.npz"))
    for case in ["default", "pcn", "vrc_nonreg", "lambda2"]:
        alpha, lam, non_reg = g[f"{case}_meta"]
        lam = int(lam) if float(lam).is_integer() else float(lam)
        x = dev(torch.from_numpy(g[f"{case}_x"])).requires_grad_()
        gt = dev(torch.from_numpy(g[f"{case}_gt"])).requires_grad_()
        w = dev(torch.from_numpy(g[f"{case}_w"]))
        loss, cd_p, cd_t, dist1, dist2, idx1, idx2 = ured.calc_dcd(x, gt, alpha=alpha, n_lambda=lam, return_raw=True, non_reg=bool(non_reg))
        assert np.array_equal(idx1.cpu().numpy(), g[f"{case}_idx1"]) and np.array_equal(idx2.cpu().numpy(), g[f"{case}_idx2"])
        assert np.array_equal(dist1.detach().cpu().numpy(), g[f"{case}_dist1"]) and np.array_equal(dist2.detach().cpu().numpy(), g[f"{case}_dist2"])
        for got, name in [(loss, "loss"), (cd_p, "cd_p"), (cd_t, "cd_t")]:
            assert np.allclose(got.detach().cpu().numpy(), g[f"{case}_{name}"], rtol=RTOL, atol=0), (case, name)
        (loss * w).sum().backward()
        assert rel_err(x.grad.cpu().numpy(), g[f"{case}_g_x_loss"]) < RTOL
        assert rel_err(gt.grad.cpu().numpy(), g[f"{case}_g_gt_loss"]) < RTOL
        x.grad = None; gt.grad = None
        cd_p2, cd_t2, f1 = ured.calc_cd(x, gt, calc_f1=True)
        ((cd_p2 + 3 * cd_t2) * w).sum().backward()
        assert np.allclose(f1.cpu().numpy(), g[f"{case}_f1"], rtol=1e-6)
        assert rel_err(x.grad.cpu().numpy(), g[f"{case}_g_x_cd"]) < RTOL
        assert rel_err(gt.grad.cpu().numpy(), g[f"{case}_g_gt_cd"]) < RTOL


def test_calc_cd_variants(ured, oracle):
    x0, gt0 = make_clouds(50, 2, 300, "S"), make_clouds(51, 2, 400, "S")
    x, gt = dev(x0), dev(gt0)
    cd_p, cd_t = ured.calc_cd(x, gt)
    op, ot = oracle.t.calc_cd_oracle(x0, gt0)
    assert np.allclose(cd_p.cpu().numpy(), op.numpy(), rtol=RTOL) and np.allclose(cd_t.cpu().numpy(), ot.numpy(), rtol=RTOL)
    sep_p, sep_t = ured.calc_cd(x, gt, separate=True)
    assert sep_p.shape == (2, 2) and np.allclose((sep_t[0] + sep_t[1]).cpu().numpy(), ot.numpy(), rtol=RTOL)
    raw = ured.calc_cd(x, gt, return_raw=True)
    assert len(raw) == 6 and raw[4].dtype == torch.int32
    # strided inputs are made contiguous like the reference does (dist_chamfer_3D.py:72-73)
    xs = dev(torch.cat([x0, x0], 2))[:, :, :3]
    d = ured.chamfer_3DDist()(xs, gt)
    e = ured.chamfer_3DDist()(x, gt)
    assert torch.equal(d[0], e[0]) and torch.equal(d[2], e[2])
    with pytest.raises(TypeError):
        ured.chamfer_3DDist()(x.double(), gt)


def test_chamfer_loss_adaptor(ured, oracle):
    src0 = make_clouds(60, 2, 3 * 1024, "S")
    tgt0 = make_clouds(61, 2, 2048, "S")
    parts0 = [[make_clouds(62 + i, 1, 100 + 37 * i, "S")[0] for i in range(3)], [make_clouds(70 + i, 1, 64 + 11 * i, "S")[0] for i in range(2)]]
    mask = torch.tensor([[1, 1, 1], [1, 1, 0]])
    src = dev(src0).requires_grad_()
    full, part = ured.compute_cm_loss(src, dev(tgt0), [[dev(p) for p in ps] for ps in parts0], dev(mask))

    def cd2(p1, p2):
        d1, d2, _, _ = oracle.t.oracle_cd(p1, p2)
        return d1.mean(1) + d2.mean(1)
    want_full = torch.stack([cd2(src0[b:b + 1, :int(mask[b].sum()) * 1024], tgt0[b:b + 1]) for b in range(2)]).mean()
    want_part = torch.stack([torch.stack([cd2(src0[b:b + 1, i * 1024:(i + 1) * 1024], parts0[b][i][None]) for i in range(len(parts0[b]))]).mean() for b in range(2)]).mean()
    assert np.isclose(full.item(), want_full.item(), rtol=RTOL) and np.isclose(part.item(), want_part.item(), rtol=RTOL)
    (full + part).backward()
    assert torch.isfinite(src.grad).all() and src.grad.abs().sum() > 0
    assert torch.allclose(ured.chamfer_distance2(dev(src0[:, :2048]), dev(tgt0)).cpu(), cd2(src0[:, :2048], tgt0), rtol=RTOL)


def test_retrieval_scores_and_rankings(ured, oracle):
    Q, K, N, M = 3, 10, 600, 512
    tg = make_clouds(80, Q, N, "S")
    cands = torch.stack([make_clouds(81 + q, K, M, "S") * (0.8 + 0.05 * q) for q in range(Q)])
    sc = ured.score_candidates(dev(tg), dev(cands), alpha=1000, n_lambda=1)
    want = {"dcd": [], "cd_p": [], "cd_t": []}
    for q in range(Q):
        l, p, t = oracle.t.calc_dcd_oracle(cands[q], tg[q:q + 1].expand(K, N, 3))
        want["dcd"].append(l); want["cd_p"].append(p); want["cd_t"].append(t)
    for key in want:
        w = torch.stack(want[key])
        assert np.allclose(sc[key].cpu().numpy(), w.numpy(), rtol=RTOL), key
        # rankings: full ascending order identical to the oracle's (score, index) order
        got_v, got_i = ured.topk_smallest(sc[key], K)
        _, want_i = oracle.t.topk_oracle(w, K)
        assert np.array_equal(got_i.cpu().numpy(), want_i.numpy()), key
    # library scoring == candidate scoring with a shared library, in one slab and in several
    libr = make_clouds(90, 37, M, "S")
    full = ured.score_library(dev(tg), dev(libr))
    slabs = ured.score_library(dev(tg), ured.PackedClouds(dev(libr)), max_pairs=16)
    ref = ured.score_candidates(dev(tg), dev(libr.unsqueeze(0).expand(Q, 37, M, 3).contiguous()))
    for key in full:
        assert torch.equal(full[key], ref[key]) and torch.equal(slabs[key], ref[key]), key
    v, i = ured.retrieve(dev(tg), dev(libr), k=10)
    _, oi = oracle.t.topk_oracle(ref["cd_t"].cpu(), 10)
    assert np.array_equal(i.cpu().numpy(), oi.numpy())


def test_topk_ties_nan_and_offset(ured, oracle):
    s = torch.tensor([[3.0, 1.0, 1.0, float("nan"), -2.0, 1.0, 0.0, float("inf")],
                      [0.0, -0.0, 5.0, 5.0, 5.0, -1.0, 2.0, 2.0]])
    v, i = ured.topk_smallest(dev(s), 6, idx_offset=100)
    assert i.cpu().tolist() == [[104, 106, 101, 102, 105, 100], [105, 101, 100, 106, 107, 102]]
    big = torch.rand(4, 20000, generator=torch.Generator().manual_seed(1)).round(decimals=3)  # many ties
    v, i = ured.topk_smallest(dev(big), 10)
    ov, oi = oracle.t.topk_oracle(big, 10)
    assert np.array_equal(i.cpu().numpy(), oi.numpy()) and np.array_equal(v.cpu().numpy(), ov.numpy())


def test_sharded_retrieval_matches_single_gpu(ured):
    """Emulate 4 ranks on one GPU: shard, local top-k with global ids, merge -- same ranking as unsharded."""
    S, M, Q, k = 50, 256, 3, 10
    libr = dev(make_clouds(95, S, M, "S"))
    tg = dev(make_clouds(96, Q, 300, "S"))
    v, i = ured.retrieve(tg, libr, k=k)
    loc_s, loc_i = [], []
    for r in range(4):
        lo, hi = ured.shard_bounds(S, 4, r)
        sc = ured.score_library(tg, libr[lo:hi])["cd_t"]
        a, b = ured.topk_smallest(sc, min(k, hi - lo), idx_offset=lo)
        loc_s.append(a); loc_i.append(b)
    ms, mi = ured.merge_topk(torch.cat(loc_s, 1), torch.cat(loc_i, 1), k)
    assert torch.equal(mi, i) and torch.equal(ms, v)
    ms, mi = ured.retrieve_sharded(tg, libr, 0, k=k)  # world size 1 path
    assert torch.equal(mi, i)


def test_stream_and_error_behaviour(ured):
    a, b = dev(make_clouds(0, 2, 400, "S")), dev(make_clouds(1, 2, 300, "S"))
    want = ured.chamfer_3DDist()(a, b)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = ured.chamfer_3DDist()(a, b)
    s.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(got, want))
    with pytest.raises(ValueError):
        ured.chamfer_3DDist()(a, b[:1])
    empty = ured.chamfer_3DDist()(a[:0], b[:0])
    assert empty[0].shape == (0, 400)
    z = ured.chamfer_3DDist()(a, b[:, :0])  # empty opposing cloud: the reference leaves its zero-filled outputs
    assert (z[0] == 0).all() and (z[2] == 0).all() and z[1].shape == (2, 0)


@pytest.mark.parametrize("alpha,lam,non_reg", [(1000, 1, False), (60, 0.5, True), (0.0, 1, False)])
def test_ragged_batches_match_per_sample_calls(ured, oracle, alpha, lam, non_reg):
    """chamfer_ragged(x, gt, len_x, len_gt) == calc_dcd on each sample's own slices (values, raw outputs, gradients)."""
    B, NX, NG = 6, 1536, 1100
    x0, gt0 = make_clouds(110, B, NX, "S"), make_clouds(111, B, NG, "S") * 0.9
    len_x = torch.tensor([1536, 1024, 1, 700, 33, 0], dtype=torch.int32)
    len_gt = torch.tensor([1100, 5, 640, 1100, 0, 77], dtype=torch.int32)
    w = torch.linspace(0.5, 1.5, B)
    x, gt = dev(x0).requires_grad_(), dev(gt0).requires_grad_()
    loss, cd_p, cd_t, d1, d2, i1, i2 = ured.chamfer_ragged(x, gt, dev(len_x), dev(len_gt), alpha=alpha, n_lambda=lam, non_reg=non_reg)
    ((loss + 2.0 * cd_t) * dev(w)).sum().backward()
    torch.cuda.synchronize()
    gx, ggt = x.grad.cpu(), gt.grad.cpu()
    for b in range(B):
        lx, lg = int(len_x[b]), int(len_gt[b])
        assert (d1[b, lg:] == 0).all() and (i1[b, lg:] == 0).all() and (d2[b, lx:] == 0).all()
        assert (gx[b, lx:] == 0).all() and (ggt[b, lg:] == 0).all()
        if lx == 0 or lg == 0:
            assert loss[b] == 0 and cd_t[b] == 0 and (d1[b] == 0).all() and (d2[b] == 0).all()
            assert (gx[b] == 0).all() and (ggt[b] == 0).all()
            continue
        xo, gto = x0[b:b + 1, :lx].clone().requires_grad_(), gt0[b:b + 1, :lg].clone().requires_grad_()
        ol, op, ot, od1, od2, oi1, oi2 = oracle.t.calc_dcd_oracle(xo, gto, alpha=alpha, n_lambda=lam, return_raw=True, non_reg=non_reg)
        assert np.array_equal(i1[b, :lg].cpu().numpy(), oi1[0].numpy()) and np.array_equal(i2[b, :lx].cpu().numpy(), oi2[0].numpy())
        assert np.array_equal(d1[b, :lg].detach().cpu().numpy(), od1[0].detach().numpy())
        assert np.array_equal(d2[b, :lx].detach().cpu().numpy(), od2[0].detach().numpy())
        for got, want in [(loss[b], ol[0]), (cd_p[b], op[0]), (cd_t[b], ot[0])]:
            assert np.isclose(got.item(), want.item(), rtol=RTOL, atol=1e-12)
        ((ol + 2.0 * ot) * w[b]).sum().backward()
        assert rel_err(gx[b, :lx].numpy(), xo.grad[0].numpy()) < RTOL
        assert rel_err(ggt[b, :lg].numpy(), gto.grad[0].numpy()) < RTOL


def test_compute_cm_loss_has_no_per_sample_launches(ured):
    """The batched compute_cm_loss issues a fixed number of device operations, independent of batch size and part count:
    counted over EVERYTHING the GPU executes (torch's kernels and copies included, via the profiler), not just this
    library's own launch counter."""
    from torch.profiler import ProfilerActivity, profile
    lib = ured._native.load()

    def run(B, P):
        src = dev(make_clouds(120, B, P * 1024, "S")).requires_grad_()
        tgt = dev(make_clouds(121, B, 2048, "S"))
        parts = [[dev(make_clouds(122 + i, 1, 50 + 13 * i, "S")[0]) for i in range(P)] for _ in range(B)]
        mask = torch.ones(B, P, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        n0 = lib.ured_kernel_launches()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            full, part = ured.compute_cm_loss(src, tgt, parts, mask)
            (full + part).backward()
            torch.cuda.synchronize()
        import collections
        device_ops = collections.Counter(e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
        return lib.ured_kernel_launches() - n0, device_ops

    run(2, 2)                                   # warm-up (lazy module loads)
    run(8, 4)
    own_a, all_a = run(2, 2)
    own_b, all_b = run(8, 4)
    assert own_a == own_b
    assert sum(all_a.values()) > 0, "the profiler saw no device activity"
    diff = {k: (all_a.get(k, 0), all_b.get(k, 0)) for k in set(all_a) | set(all_b) if all_a.get(k, 0) != all_b.get(k, 0)}
    assert sum(all_a.values()) == sum(all_b.values()), f"device operations grow with the batch (B=2,P=2 vs B=8,P=4): {diff}"


def test_backward_is_bit_reproducible_and_handles_crowded_points(ured, oracle):
    """grad_gather_kernel: no float atomics -> the same bits on every run; long 'chosen-by' lists (many points sharing one
    nearest neighbour -- the density pathology DCD measures) take the linear-scan path; both agree with the float64 oracle
    and with the general global-atomic kernels."""
    import os
    B, N, M = 3, 2048, 2048
    a = make_clouds(180, B, N, "S")
    b = make_clouds(181, B, M, "S") * 3.0 + 5.0           # far away cloud: every point of a chooses one of very few points of b
    b[:, :40] = make_clouds(182, B, 40, "S") * 0.5         # ... a small cluster inside a: lists of ~50 entries each
    g = torch.Generator().manual_seed(6)
    w1, w2 = torch.randn(B, N, generator=g), torch.randn(B, M, generator=g)

    def grads():
        xa, xb = dev(a).requires_grad_(), dev(b).requires_grad_()
        d1, d2, i1, i2 = ured.chamfer_3DDist()(xa, xb)
        ((d1 * dev(w1)).sum() + (d2 * dev(w2)).sum()).backward()
        return xa.grad.clone(), xb.grad.clone(), i1, i2

    g1, g2, i1, i2 = grads()
    counts = torch.bincount(i1[0].long(), minlength=M)
    assert counts.max().item() > 32, "the test should exercise lists longer than the in-place sort limit"
    for _ in range(3):
        h1, h2, _, _ = grads()
        assert torch.equal(g1, h1) and torch.equal(g2, h2), "backward is not bit-reproducible"
    o = oracle.c.chamfer_forward(a.numpy(), b.numpy())
    r1, r2 = oracle.c.chamfer_backward_f64(a.numpy(), b.numpy(), w1.numpy(), w2.numpy(), o[2], o[3])
    assert rel_err(g1.cpu().numpy(), r1) < RTOL and rel_err(g2.cpu().numpy(), r2) < RTOL
    os.environ["URED_GRAD_GENERAL"] = "1"                  # the general kernels (global atomics) on the same problem
    try:
        k1, k2, _, _ = grads()
    finally:
        del os.environ["URED_GRAD_GENERAL"]
    assert rel_err(k1.cpu().numpy(), r1) < RTOL and rel_err(k2.cpu().numpy(), r2) < RTOL


@pytest.mark.parametrize("B,N,M,ragged", [(3, 2048, 2048, False), (5, 1500, 700, True), (2, 6000, 8000, False), (80, 512, 512, False)])
def test_gather_backward_kernels_agree_bit_for_bit(ured, B, N, M, ragged):
    """The three atomic-free backward kernels -- one CTA per pair (grad_gather_kernel), one CTA per (pair, cloud) narrow and
    wide (grad_side_kernel<256/1024>) -- subtract in the same order (own term, then the choosers by ascending index), so they
    must return the same bits; the library picks between them by batch size and cloud size (ured_dcd_backward)."""
    import os
    a, b = make_clouds(190, B, N, "S"), make_clouds(191, B, M, "S") * 0.9
    g = torch.Generator().manual_seed(8)
    w1, w2 = torch.randn(B, N, generator=g), torch.randn(B, M, generator=g)
    lens = {}
    if ragged:
        lens = dict(len_x=torch.randint(1, N + 1, (B,), generator=g).int().cuda(), len_gt=torch.randint(1, M + 1, (B,), generator=g).int().cuda())
        lens["len_x"][0] = 0

    def grads(env):
        for k in ("URED_GRAD_PAIR_CTA", "URED_GRAD_SIDE_WIDE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            xa, xb = dev(a).requires_grad_(), dev(b).requires_grad_()
            if ragged:
                loss, cd_p, cd_t = ured.chamfer_ragged(xa, xb, alpha=200, n_lambda=0.5, **lens)[:3]
                (loss.sum() + 0.3 * cd_t.sum() + 0.1 * cd_p.sum()).backward()
            else:
                d1, d2, _, _ = ured.chamfer_3DDist()(xa, xb)
                ((d1 * dev(w1)).sum() + (d2 * dev(w2)).sum()).backward()
            return xa.grad.clone(), xb.grad.clone()
        finally:
            for k in env:
                os.environ.pop(k, None)

    default = grads({})
    fits_pair = (N + M) * 18 <= 200 * 1024
    variants = [{"URED_GRAD_PAIR_CTA": "0", "URED_GRAD_SIDE_WIDE": "0"}, {"URED_GRAD_PAIR_CTA": "0", "URED_GRAD_SIDE_WIDE": "1"}]
    if fits_pair:
        variants.append({"URED_GRAD_PAIR_CTA": "1"})
    for env in variants:
        got = grads(env)
        assert torch.equal(got[0], default[0]) and torch.equal(got[1], default[1]), f"{env} differs from the default kernel"
    general = grads({"URED_GRAD_GENERAL": "1"})                 # the global-atomic kernels: same values up to summation order
    assert rel_err(general[0].cpu().numpy(), default[0].cpu().numpy()) < RTOL
    assert rel_err(general[1].cpu().numpy(), default[1].cpu().numpy()) < RTOL


def test_knn1_and_residual_retrieval_loss(ured, oracle):
    """K=1 kNN (loss/basic_loss.py:249-265's knn_points call) on the one-direction NN kernel, ragged source lengths."""
    B, N, P = 3, 700, 3
    x0, src0 = make_clouds(130, B, N, "S"), make_clouds(131, B, P * 1024, "S")
    res0 = 0.01 * torch.randn(B, N, 3, generator=torch.Generator().manual_seed(3))
    mask = torch.tensor([[1, 1, 1], [1, 0, 0], [1, 1, 0]])
    x, src, res = dev(x0).requires_grad_(), dev(src0).requires_grad_(), dev(res0).requires_grad_()
    loss, reg = ured.residual_retrieval_loss(x, src, res, dev(mask))
    (loss + reg).backward()
    xo, so, ro = x0.clone().requires_grad_(), src0.clone().requires_grad_(), res0.clone().requires_grad_()
    nn_all = []
    for b in range(B):
        cnt = int(mask[b].sum()) * 1024
        d1, _, i1, _ = oracle.t.oracle_cd(xo[b:b + 1], so[b:b + 1, :cnt])
        nn_all.append(so[b, :cnt][i1[0].long()])
        dists, idx, nn = ured.knn1_points(dev(x0[b:b + 1]), dev(src0[b:b + 1, :cnt]))
        assert np.array_equal(idx[0, :, 0].cpu().numpy(), i1[0].numpy().astype(np.int64))
        assert np.array_equal(dists[0, :, 0].cpu().numpy(), d1[0].detach().numpy())
    want = torch.mean(torch.sum(torch.abs(xo + ro - torch.stack(nn_all)), dim=-1))
    want_reg = torch.mean(torch.sum(torch.abs(ro), dim=-1))
    (want + want_reg).backward()
    assert np.isclose(loss.item(), want.item(), rtol=RTOL) and np.isclose(reg.item(), want_reg.item(), rtol=RTOL)
    assert rel_err(src.grad.cpu().numpy(), so.grad.numpy()) < RTOL and rel_err(x.grad.cpu().numpy(), xo.grad.numpy()) < RTOL
    # distances are differentiable in both clouds
    a, b_ = dev(x0).requires_grad_(), dev(src0[:, :1024]).requires_grad_()
    d, _, _ = ured.knn1_points(a, b_, return_nn=False)
    d.sum().backward()
    e1, _, j1, _ = oracle.c.chamfer_forward(x0.numpy(), src0[:, :1024].numpy())
    r1, r2 = oracle.c.chamfer_backward_f64(x0.numpy(), src0[:, :1024].numpy(), np.ones((B, N), np.float32), np.zeros((B, 1024), np.float32), j1,
                                           np.zeros((B, 1024), np.int32))
    assert rel_err(a.grad.cpu().numpy(), r1) < RTOL and rel_err(b_.grad.cpu().numpy(), r2) < RTOL


def test_all_pairs_and_pickle_layout(ured, oracle, tmp_path):
    """score_all_pairs == the reference's get_src_pair rows (engine/generate_pair.py:69-85), pickles readable its way."""
    import pickle
    S, M = 9, 400
    libr = make_clouds(140, S, M, "S") * torch.linspace(0.7, 1.1, S).view(S, 1, 1)
    sc = ured.score_all_pairs(dev(libr), max_pairs=20)
    names = [f"shape{i}" for i in range(S)]
    paths = ured.write_pair_pickles(str(tmp_path), names, sc)
    for idx in [0, 4, 8]:
        rec = pickle.load(open(paths[idx], "rb"))
        assert set(rec) == {"dcd_loss", "cd_s", "cd_m"} and rec["cd_m"].shape == (S - idx,) and rec["cd_m"].dtype == np.float64
        l, p, t = oracle.t.calc_dcd_oracle(libr[idx:], libr[idx:idx + 1].expand(S - idx, M, 3))
        assert np.allclose(rec["dcd_loss"], l.numpy(), rtol=RTOL) and np.allclose(rec["cd_s"], p.numpy(), rtol=RTOL)
        assert np.allclose(rec["cd_m"], t.numpy(), rtol=RTOL)
        # dataset_utils.read_pickle_topk: torch.topk(torch.tensor(data['cd_m']), k, largest=False)
        k = min(3, S - idx)
        want = torch.sort(t, stable=True).indices[:k]
        got = ured.topk_smallest(dev(torch.tensor(rec["cd_m"], dtype=torch.float32)), k)[1]
        assert got.cpu().tolist() == want.tolist()
    assert (sc[:, 5, :5] == 0).all()


def test_retrieval_engine_graph_matches_eager(ured):
    S, M, Q, k = 40, 256, 3, 10
    libr = dev(make_clouds(150, S, M, "S"))
    want_v, want_i = ured.retrieve(dev(make_clouds(151, Q, 300, "S")), libr, k=k)
    eng = ured.RetrievalEngine(libr, 0, Q, k=k)
    for seed in (151, 152, 151):  # replay with new inputs, then the first again
        tg = dev(make_clouds(seed, Q, 300, "S"))
        v, i = eng.query(tg)
        ev, ei = ured.retrieve(tg, libr, k=k)
        assert torch.equal(i, ei) and torch.equal(v, ev)
    assert torch.equal(i, want_i) and torch.equal(v, want_v)
    small = ured.RetrievalEngine(libr[:4], 100, Q, k=k, use_graph=False)  # shard shorter than k: padded
    v, i = small.query(dev(make_clouds(151, Q, 300, "S")))
    assert (i[:, 4:] == -1).all() and (i[:, :4] >= 100).all() and torch.isinf(v[:, 4:]).all()


def test_full_size_properties_cfg2(ured):
    """BASELINE configs[1] at full size (640 pairs of 2048 x 2048): properties that need no CPU oracle."""
    B, N = 640, 2048
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, N, 3, generator=g); x = (x - x.mean(1, keepdim=True)); x = (x / x.norm(dim=2).amax(1).view(B, 1, 1)).cuda()
    y = (x[torch.randperm(B, generator=g)] * 0.97 + 0.01 * torch.randn(B, N, 3, generator=g).cuda()).contiguous()
    d1, d2, i1, i2 = ured.nn_forward(x, y)
    # (1) the two kernels agree bit for bit
    e = ured.nn_forward(x, y, exact_only=True)
    assert all(torch.equal(a, b) for a, b in zip((d1, d2, i1, i2), e))
    # (2) swapping the clouds swaps the outputs (the squared difference is exactly antisymmetric-invariant)
    s1, s2, j1, j2 = ured.nn_forward(y, x)
    assert torch.equal(s1, d2) and torch.equal(s2, d1) and torch.equal(j1, i2) and torch.equal(j2, i1)
    # (3) the reported distance is the distance to the reported neighbour (same difference form, fp32 round-off only)
    nb = torch.gather(y, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    chk = ((nb - x) ** 2).sum(-1)
    assert torch.allclose(chk, d1, rtol=1e-5, atol=1e-12)
    # (4) optimality on a sample of queries against a float64 brute force
    qs = torch.randint(0, N, (64,), generator=g).cuda()
    for b in (0, 319, 639):
        full = ((x[b, qs].double().unsqueeze(1) - y[b].double().unsqueeze(0)) ** 2).sum(-1)
        assert torch.allclose(full.min(1).values.float(), d1[b, qs], rtol=1e-5, atol=1e-12)
    # (5) permuting the candidates permutes the indices and leaves the distances untouched
    perm = torch.randperm(N, generator=g).cuda()
    p1, _, k1, _ = ured.nn_forward(x[:8], y[:8, perm].contiguous())
    assert torch.equal(p1, d1[:8])
    same = perm[k1.long()] == i1[:8]
    tie_ok = torch.gather(y[:8], 1, perm[k1.long()].unsqueeze(-1).expand(-1, -1, 3))
    assert torch.allclose(((tie_ok - x[:8]) ** 2).sum(-1)[~same], chk[:8][~same], rtol=1e-6, atol=0)  # differing picks are ties
    # (6) idempotence of the DCD scores and rankings across repeated calls (deterministic forward)
    a = ured.calc_dcd(y, x)
    b2 = ured.calc_dcd(y, x)
    assert all(torch.equal(u, v) for u, v in zip(a, b2))


def test_backward_general_path_large_pairs(ured, oracle):
    """Pairs too large for shared memory (18 bytes per point > 200 KB) take grad_kernel<own/scatter> (global atomics)."""
    B, N, M = 2, 7000, 6000
    a, b = make_clouds(160, B, N, "S"), make_clouds(161, B, M, "S")
    g = torch.Generator().manual_seed(4)
    w1, w2 = torch.randn(B, N, generator=g), torch.randn(B, M, generator=g)
    xa, xb = dev(a).requires_grad_(), dev(b).requires_grad_()
    d1, d2, i1, i2 = ured.chamfer_3DDist()(xa, xb)
    ((d1 * dev(w1)).sum() + (d2 * dev(w2)).sum()).backward()
    o = oracle.c.chamfer_forward(a.numpy(), b.numpy())
    assert np.array_equal(i1.cpu().numpy(), o[2]) and np.array_equal(i2.cpu().numpy(), o[3])
    r1, r2 = oracle.c.chamfer_backward_f64(a.numpy(), b.numpy(), w1.numpy(), w2.numpy(), o[2], o[3])
    assert rel_err(xa.grad.cpu().numpy(), r1) < RTOL and rel_err(xb.grad.cpu().numpy(), r2) < RTOL
    # ragged lengths on the same path, DCD loss
    len_x, len_gt = torch.tensor([4500, 5000], dtype=torch.int32), torch.tensor([6000, 3333], dtype=torch.int32)
    x, gt = dev(b).requires_grad_(), dev(a).requires_grad_()
    loss = ured.chamfer_ragged(x, gt, dev(len_x), dev(len_gt), alpha=200, n_lambda=0.5)[0]
    loss.sum().backward()
    for s in range(B):
        lx, lg = int(len_x[s]), int(len_gt[s])
        xo, gto = b[s:s + 1, :lx].clone().requires_grad_(), a[s:s + 1, :lg].clone().requires_grad_()
        ol = oracle.t.calc_dcd_oracle(xo, gto, alpha=200, n_lambda=0.5)[0]
        ol.sum().backward()
        assert np.isclose(loss[s].item(), ol.item(), rtol=RTOL)
        assert rel_err(x.grad[s, :lx].cpu().numpy(), xo.grad[0].numpy()) < RTOL
        assert rel_err(gt.grad[s, :lg].cpu().numpy(), gto.grad[0].numpy()) < RTOL
        assert (x.grad[s, lx:] == 0).all() and (gt.grad[s, lg:] == 0).all()


def test_knn1_with_candidate_splits(ured, oracle):
    """One-direction search on a shape that is cut into candidate splits (few pairs, large clouds)."""
    p1, p2 = make_clouds(170, 1, 2048, "S"), make_clouds(171, 1, 4100, "S")
    lib = ured._native.load()
    assert lib.ured_nn_scratch_bytes(1, 2048, 4100) > 0
    dists, idx, nn = ured.knn1_points(dev(p1), dev(p2))
    d1, _, i1, _ = oracle.c.chamfer_forward(p1.numpy(), p2.numpy())
    assert np.array_equal(idx[0, :, 0].cpu().numpy(), i1[0].astype(np.int64))
    assert np.array_equal(dists[0, :, 0].cpu().numpy(), d1[0])
    assert torch.equal(nn[0, :, 0].cpu(), p2[0][torch.from_numpy(i1[0]).long()])
    # ragged candidates on the split path
    dists, idx, _ = ured.knn1_points(dev(p1), dev(p2), lengths2=torch.tensor([3000]))
    d1, _, i1, _ = oracle.c.chamfer_forward(p1.numpy(), p2[:, :3000].numpy())
    assert np.array_equal(idx[0, :, 0].cpu().numpy(), i1[0].astype(np.int64)) and np.array_equal(dists[0, :, 0].cpu().numpy(), d1[0])


def test_compat_pytorch3d_subset(ured, oracle):
    """The pytorch3d-shaped entry points the reference calls (loss/chamfer_loss.py:1, loss/basic_loss.py:257)."""
    x0, y0 = make_clouds(180, 3, 500, "S"), make_clouds(181, 3, 640, "S")
    d1, d2, i1, _ = oracle.c.chamfer_forward(x0.numpy(), y0.numpy())
    per_sample = d1.astype(np.float64).mean(1) + d2.astype(np.float64).mean(1)
    loss, normals = ured.compat.chamfer_distance(dev(x0), dev(y0), batch_reduction=None)
    assert normals is None and np.allclose(loss.cpu().numpy(), per_sample, rtol=RTOL)
    loss, _ = ured.compat.chamfer_distance(dev(x0), dev(y0))
    assert np.isclose(loss.item(), per_sample.mean(), rtol=RTOL)
    knn = ured.compat.knn_points(dev(x0), dev(y0), K=1, return_nn=True)
    assert knn.dists.shape == (3, 500, 1) and knn.idx.dtype == torch.int64 and knn.knn.shape == (3, 500, 1, 3)
    assert np.array_equal(knn.idx[..., 0].cpu().numpy(), i1.astype(np.int64))
    with pytest.raises(NotImplementedError):
        ured.compat.knn_points(dev(x0), dev(y0), K=3)


def test_non_finite_inputs_policy(ured, oracle, monkeypatch):
    """Parity is defined for finite clouds.  With a NaN / inf coordinate the call must stay memory-safe (indices in range, other
    pairs of the batch untouched), and URED_CHECK_FINITE=1 turns it into an error."""
    a, b = make_clouds(190, 3, 700, "S"), make_clouds(191, 3, 900, "S")
    bad_a, bad_b = a.clone(), b.clone()
    bad_a[1, 5, 0] = float("nan")
    bad_b[1, 512, 2] = float("inf")
    for exact_only in (False, True):
        d1, d2, i1, i2 = ured.nn_forward(dev(bad_a), dev(bad_b), exact_only=exact_only)
        assert int(i1.min()) >= 0 and int(i1.max()) < 900 and int(i2.min()) >= 0 and int(i2.max()) < 700
        o = oracle.c.chamfer_forward(a.numpy(), b.numpy())
        for s in (0, 2):                                            # the clean pairs of the batch keep their exact results
            assert np.array_equal(i1[s].cpu().numpy(), o[2][s]) and np.array_equal(d1[s].cpu().numpy(), o[0][s])
            assert np.array_equal(i2[s].cpu().numpy(), o[3][s]) and np.array_equal(d2[s].cpu().numpy(), o[1][s])
    monkeypatch.setenv("URED_CHECK_FINITE", "1")
    with pytest.raises(ValueError, match="non-finite"):
        ured.nn_forward(dev(bad_a), dev(bad_b))
    ured.nn_forward(dev(a), dev(b))                                 # finite clouds pass the check


def test_size_limits_at_the_boundary(ured, oracle):
    """dcd_fwd holds one pair's histograms in shared memory: exactly 51 200 points per pair work, one more is URED_E_RANGE;
    the general backward takes more than 65 535 pairs per call (it used to put the pairs on gridDim.y)."""
    lib = ured._native.load()
    n1, n2 = 25600, 25600
    x, gt = make_clouds(195, 1, n2, "S"), make_clouds(196, 1, n1, "S")
    res = ured.calc_dcd(dev(x), dev(gt), alpha=50, n_lambda=1)
    ref = oracle.t.calc_dcd_oracle(x, gt, alpha=50, n_lambda=1)
    for got, want in zip(res, ref):
        assert np.allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=0)
    with pytest.raises(ured.NativeLibraryError, match="51200"):
        ured.calc_dcd(dev(make_clouds(197, 1, n2 + 1, "S")), dev(gt))
    # 70 000 tiny pairs, one target broadcast over all of them -> general backward kernels (cloud 1 is shared)
    B = 70000
    tgt, cands = dev(make_clouds(198, 1, 8, "S")), dev(make_clouds(199, B, 8, "S"))
    d1, d2, i1, i2 = ured.retrieval.nn_pairs(tgt, cands, B, B, B)
    g1, g2 = torch.empty(1, 8, 3, device="cuda"), torch.empty(B, 8, 3, device="cuda")
    w1, w2 = torch.ones(B, 8, device="cuda"), torch.ones(B, 8, device="cuda")
    rc = lib.ured_chamfer_backward(tgt.data_ptr(), cands.data_ptr(), B, 8, 8, B, B, None, None, w1.data_ptr(), w2.data_ptr(),
                                   i1.data_ptr(), i2.data_ptr(), g1.data_ptr(), g2.data_ptr(), None)
    ured._native.check(rc, "ured_chamfer_backward")
    torch.cuda.synchronize()
    t64, c64 = tgt.double(), cands.double()
    diff2 = 2 * (c64 - t64[0][i2.long()])                                    # own terms of the candidate points
    chosen = torch.gather(c64, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))  # [B, 8, 3]: the candidate each target point chose
    diff1 = 2 * (t64[0].unsqueeze(0) - chosen)                               # own terms of the (broadcast) target points, per pair
    want2 = diff2.clone()
    want2.scatter_add_(1, i1.long().unsqueeze(-1).expand(-1, -1, 3), -diff1)  # ... scattered onto the chosen candidates
    assert torch.allclose(g2.double(), want2, rtol=1e-5, atol=1e-6)
    want1 = diff1.sum(0)
    want1.index_add_(0, i2.long().view(-1), (-diff2).view(-1, 3))
    assert torch.allclose(g1[0].double(), want1, rtol=1e-4, atol=1e-3)


def test_tail_split_launch_plan_returns_the_same_bits(ured, monkeypatch):
    """URED_NN_TAIL_SPLIT=1 cuts only the last partial wave of work items into candidate ranges (an experiment knob, off by
    default because it does not pay): the merged result must be bit-identical to the unsplit launch."""
    import ctypes
    lib = ured._native.load()
    B, n = 125, 2048
    a, b = dev(make_clouds(210, B, n, "S")), dev(make_clouds(211, B, n, "S") * 0.97)
    want = ured.nn_forward(a, b)
    monkeypatch.setenv("URED_NN_TC", "0")            # the tail rule belongs to the FP32-pipe kernel
    monkeypatch.setenv("URED_NN_TAIL_SPLIT", "1")
    v, q, t, ns, items, split = (ctypes.c_int() for _ in range(6))
    lib.ured_nn_launch_shape(B, n, n, 0, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns), ctypes.byref(items), ctypes.byref(split))
    assert (items.value, split.value, ns.value) == (1000, 260, 4)
    got = ured.nn_forward(a, b)
    for g, w in zip(got, want):
        assert torch.equal(g, w)
