"""GPU: retrieval rankings pinned at BASELINE scale (north_star: "idx1/idx2 and retrieval rankings bit-exact").

Reference scores are produced the reference's way -- its UNMODIFIED CUDA op (oracle/_ref) followed by its torch ops
(model_utils.py:13-58, restated in oracle/torch_path.py) -- and ranked with torch.sort(stable=True), i.e. ascending
(score, index).  Two product paths are held to those ids:
  * the fused epilogue kernel, whose float32 row sums follow torch's reduction order (include/ured_chamfer.h), and
  * exact_ranking=True, which calls torch's own reductions on the bit-exact dist/idx.
The row-sum emulation itself is pinned against torch.mean bit for bit (first test): if a torch release ever changes
its reduction schedule that test says so, and exact_ranking=True remains correct by construction.
"""
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_op(oracle):
    mod = oracle.build.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref not built (needs /root/reference in the authoring container)")
    return mod


def deformed(seed, targets, K):
    """K candidates per target: anisotropically scaled, jittered resamples (bench.py's synthetic candidates)."""
    g = torch.Generator().manual_seed(seed)
    Q, n, _ = targets.shape
    base = targets.repeat_interleave(K, dim=0)
    perm = torch.stack([torch.randperm(n, generator=g) for _ in range(Q * K)])
    x = torch.gather(base, 1, perm.unsqueeze(-1).expand(-1, -1, 3))
    return x * (1 + 0.1 * torch.rand(Q * K, 1, 3, generator=g)) + 0.01 * torch.randn(Q * K, n, 3, generator=g)


@pytest.mark.parametrize("B,n1,n2", [(16, 2048, 2048), (640, 2048, 2048), (37, 2000, 1000), (16, 4100, 2047), (33, 1026, 130),
                                     (1000, 2048, 2048), (16, 200, 8000)])
def test_fused_row_means_carry_torch_bits(ured, B, n1, n2):
    """cd_t / cd_p of the epilogue kernel == torch's mean(1) of the same rows, bit for bit (128 < n < 8192, rows >= 16)."""
    lib = ured._native.load()
    g = torch.Generator().manual_seed(B + n1)
    d1 = (torch.rand(B, n1, generator=g) ** 4 * 0.01).cuda()
    d2 = (torch.rand(B, n2, generator=g) ** 4 * 0.01).cuda()
    i1 = torch.randint(0, n2, (B, n1), generator=g).int().cuda()
    i2 = torch.randint(0, n1, (B, n2), generator=g).int().cuda()
    _, cd_p, cd_t = ured.retrieval.pair_scores(d1, d2, i1, i2, alpha=1000, n_lambda=1)
    want_t = d1.mean(1) + d2.mean(1)
    want_p = (torch.sqrt(d1).mean(1) + torch.sqrt(d2).mean(1)) / 2
    assert torch.equal(cd_t, want_t), f"{(cd_t != want_t).sum().item()} of {B} cd_t values differ in the last bits"
    assert torch.equal(cd_p, want_p)


@pytest.mark.parametrize("alpha,lam", [(1000, 1), (200, 0.5), (40, 0.5), (50, 2)])
def test_fused_dcd_loss_carries_torch_bits(ured, alpha, lam):
    B, n = 48, 2048
    x, gt = make_clouds(3, B, n, "S").cuda(), (make_clouds(4, B, n, "S") * 0.97).cuda()
    raw = ured.nn_forward(gt, x)
    fused = ured.retrieval.pair_scores(*raw, alpha=alpha, n_lambda=lam)
    torch_ops = ured.retrieval.pair_scores(*raw, alpha=alpha, n_lambda=lam, exact_ranking=True)
    for name, a, b in zip(("dcd", "cd_p", "cd_t"), fused, torch_ops):
        assert torch.equal(a, b), f"{name}: {(a != b).sum().item()} of {B} values differ"


def _reference_scores(oracle, ref_op, x, gt):
    from oracle import ref_cuda
    return ref_cuda.scores(ref_op, x, gt, alpha=1000, n_lambda=1)   # (dcd, cd_p, cd_t)


def test_rankings_cfg2_scale(ured, oracle, ref_op):
    """BASELINE configs[1]: 64 targets x K=10 deformed candidates x 2048 points -- every metric's full ranking."""
    Q, K, n = 64, 10, 2048
    targets = make_clouds(21, Q, n, "S")
    cands = deformed(22, targets, K)
    ref = _reference_scores(oracle, ref_op, cands.cuda(), targets.repeat_interleave(K, dim=0).cuda())
    for exact in (False, True):
        got = ured.score_candidates(targets.cuda(), cands.view(Q, K, n, 3).cuda(), exact_ranking=exact)
        for key, r in zip(("dcd", "cd_p", "cd_t"), ref):
            want_ids = torch.sort(r.view(Q, K), dim=1, stable=True).indices.int()
            ids = ured.topk_smallest(got[key], K)[1]
            assert torch.equal(ids, want_ids), f"{key} ranking differs (exact_ranking={exact})"
            assert torch.equal(got[key].view(-1), r), f"{key} scores are not the reference's bits (exact_ranking={exact})"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rankings_cfg3_scale(ured, oracle, ref_op, seed):
    """BASELINE configs[2]: one target against a 1000-shape library, top-10 and the full order."""
    S, n, k = 1000, 2048, 10
    lib_x = make_clouds(100 + seed, S, n, "S") * (1 + 0.05 * torch.rand(S, 1, 3, generator=torch.Generator().manual_seed(seed)))
    target = make_clouds(200 + seed, 1, n, "S")
    ref = _reference_scores(oracle, ref_op, lib_x.cuda(), target.expand(S, -1, -1).contiguous().cuda())
    want_sorted = torch.sort(ref[2], stable=True).indices.int()
    for exact in (False, True):
        sc = ured.score_library(target.cuda(), lib_x.cuda(), exact_ranking=exact)["cd_t"]
        assert torch.equal(sc.view(-1), ref[2])
        assert torch.equal(ured.topk_smallest(sc, k)[1].view(-1), want_sorted[:k])
        assert torch.equal(ured.topk_smallest(sc, S)[1].view(-1)[:1000], want_sorted)
    # the engine (graph replay, resident packed library) returns the same ids
    eng = ured.RetrievalEngine(lib_x.cuda(), 0, 1, k=k)
    assert torch.equal(eng.query(target.cuda())[1].view(-1), want_sorted[:k])


def test_rankings_10k_library(ured, oracle, ref_op):
    """BASELINE configs[4] sweep point: S = 10 000 shapes x 2048 points (scored here in slabs, by the reference in one batch)."""
    S, n, k = 10000, 2048, 10
    g = torch.Generator().manual_seed(9)
    base = make_clouds(300, 200, n, "S")
    lib_x = (base.repeat(S // 200, 1, 1) * (1 + 0.1 * torch.rand(S, 1, 3, generator=g)) + 0.003 * torch.randn(S, n, 3, generator=g)).contiguous()
    target = make_clouds(301, 1, n, "S")
    lib_dev = lib_x.cuda()
    ref = _reference_scores(oracle, ref_op, lib_dev, target.expand(S, -1, -1).contiguous().cuda())
    want_sorted = torch.sort(ref[2], stable=True).indices.int()
    sc = ured.score_library(target.cuda(), lib_dev)["cd_t"]
    assert torch.equal(sc.view(-1), ref[2])
    assert torch.equal(ured.topk_smallest(sc, k)[1].view(-1), want_sorted[:k])
    assert torch.equal(ured.topk_smallest(sc, 1024)[1].view(-1), want_sorted[:1024])


def test_fused_fscore_matches_the_reference_function(ured):
    """metrics/CD/fscore.py:3-16 inside the epilogue kernel (fscore_fused, calc_cd(calc_f1=True)) == the four torch ops, bit for bit:
    the counts are exact in any order and the factor is torch's."""
    g = torch.Generator().manual_seed(12)
    for B, n1, n2, thr in [(16, 2048, 2048, 1e-4), (37, 2000, 1000, 1e-3), (3, 300, 500, 1e-4)]:
        d1 = (torch.rand(B, n1, generator=g) ** 6 * 0.01).cuda()
        d2 = (torch.rand(B, n2, generator=g) ** 6 * 0.01).cuda()
        f, p1, p2 = ured.fscore_fused(d1, d2, thr)
        wf, wp1, wp2 = ured.fscore(d1, d2, thr)
        assert torch.equal(p1, wp1) and torch.equal(p2, wp2) and torch.equal(f, wf)
    d1[0] = 1.0; d2[0] = 1.0                      # no point under the threshold: 0/0 is reported as 0 (fscore.py:15)
    f, _, _ = ured.fscore_fused(d1, d2, 1e-4)
    assert f[0].item() == 0.0
    x, gt = make_clouds(5, 4, 1024, "S").cuda(), (make_clouds(6, 4, 1024, "S") * 0.99).cuda()
    cd_p, cd_t, f1 = ured.calc_cd(x, gt, calc_f1=True)
    dist1, dist2, _, _ = ured.chamfer_3DDist()(gt, x)
    assert torch.equal(f1, ured.fscore(dist1, dist2)[0])


def test_engine_metric_selection_and_exact_ranking(ured, oracle, ref_op):
    """score_library(metrics=...) returns only what was asked for (cd_t alone skips the DCD histograms) with unchanged bits; the
    engine in exact_ranking mode (the reference's torch reductions) returns the same ids as the fused default."""
    S, n, k = 300, 2048, 10
    lib_x = make_clouds(400, S, n, "S").cuda()
    tg = make_clouds(401, 2, n, "S").cuda()
    full = ured.score_library(tg, lib_x)
    only = ured.score_library(tg, lib_x, metrics=("cd_t",))
    assert set(only) == {"cd_t"} and torch.equal(only["cd_t"], full["cd_t"])
    two = ured.score_library(tg, lib_x, metrics=("dcd", "cd_p"))
    assert set(two) == {"dcd", "cd_p"} and torch.equal(two["dcd"], full["dcd"]) and torch.equal(two["cd_p"], full["cd_p"])
    fused = ured.RetrievalEngine(lib_x, 0, 2, k=k).query(tg)
    exact = ured.RetrievalEngine(lib_x, 0, 2, k=k, exact_ranking=True, use_graph=False).query(tg)
    assert torch.equal(fused[1], exact[1]) and torch.equal(fused[0], exact[0])
    # results survive the next query (the graph's static outputs are cloned)
    again = ured.RetrievalEngine(lib_x, 0, 2, k=k)
    first = again.query(tg)
    keep = (first[0].clone(), first[1].clone())
    again.query(make_clouds(402, 2, n, "S").cuda())
    assert torch.equal(first[0], keep[0]) and torch.equal(first[1], keep[1])


def test_pipelined_engine_equals_one_at_a_time(ured):
    """pipeline_depth=2: batches submitted back to back on alternating lanes return exactly what query() returns for each batch,
    in submission order, also when the results are picked up late and when they are copied to pinned host memory by the lane."""
    S, n, k, Q = 400, 2048, 10, 2
    lib_x = make_clouds(410, S, n, "S").cuda()
    batches = [make_clouds(420 + j, Q, n, "S").cuda() for j in range(7)]
    plain = ured.RetrievalEngine(lib_x, 0, Q, k=k)
    want = [tuple(t.clone() for t in plain.query(b)) for b in batches]
    piped = ured.RetrievalEngine(lib_x, 0, Q, k=k, pipeline_depth=2)
    assert len(piped.lanes) == 2 and piped.lanes[0].lib is piped.lib       # the library is shared, not copied
    host = [(torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int32).pin_memory()) for _ in batches]
    pend = [piped.submit(b, host_out=h) for b, h in zip(batches, host)]
    piped.drain()
    torch.cuda.synchronize()
    for j, (p, w) in enumerate(zip(pend, want)):
        got = p.result()
        assert torch.equal(got[0], w[0]) and torch.equal(got[1], w[1]), f"batch {j}"
        assert torch.equal(host[j][0], w[0].cpu()) and torch.equal(host[j][1], w[1].cpu()), f"batch {j} (host copy)"
    for b, w in zip(batches[:3], want):                                   # query() on a pipelined engine: one at a time
        got = piped.query(b)
        assert torch.equal(got[1], w[1])
    piped.check()
    piped.close()
