"""CPU: host-side logic of the retrieval path -- sharding, top-k merge, and the world_size-2
all_gather exchange over gloo (the same code path runs over NCCL on the GPU box)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


from helpers_merge import cpu_merge  # noqa: E402  (tests/helpers_merge.py: the reference merge used by the CPU tests)


def test_shard_bounds_cover_library(ured):
    for S in [0, 1, 7, 8, 125, 1000, 100003]:
        for world in [1, 2, 4, 8]:
            spans = [ured.shard_bounds(S, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (S + world - 1) // world


def test_product_merge_has_no_cpu_path(ured):
    with pytest.raises(RuntimeError, match="GPU tensors only"):
        ured.merge_topk(torch.zeros(1, 3), torch.zeros(1, 3, dtype=torch.int32), 2)


def test_reference_merge_is_lexicographic(ured, oracle):
    g = torch.Generator().manual_seed(0)
    scores = torch.rand(5, 1000, generator=g).round(decimals=2)  # heavy ties
    k = 10
    cols = []
    for r in range(8):
        lo, hi = ured.shard_bounds(1000, 8, r)
        s, i = oracle.t.topk_oracle(scores[:, lo:hi], k)
        cols.append((s, i + lo))
    ms, mi = cpu_merge(torch.cat([c[0] for c in cols], 1), torch.cat([c[1] for c in cols], 1), k)
    ws, wi = oracle.t.topk_oracle(scores, k)
    assert torch.equal(mi, wi) and torch.equal(ms, ws)
    # padding entries (id -1) never surface
    s = torch.tensor([[0.5, 0.1, 9.0]]); i = torch.tensor([[4, 2, -1]], dtype=torch.int32)
    ms, mi = cpu_merge(s, i, 3)
    assert mi.tolist() == [[2, 4, -1]] and ms[0, 2] == float("inf")


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ured_b200 as ured
    from oracle import torch_path
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers_merge import cpu_merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        scores = torch.rand(4, 37, generator=g).round(decimals=1)  # same on every rank; ties across shards
        k = 10
        lo, hi = ured.shard_bounds(37, world, rank)
        ls, li = torch_path.topk_oracle(scores[:, lo:hi], min(k, hi - lo))
        ms, mi = ured.gather_and_merge(ls, (li + lo).int(), k, merge=cpu_merge)
        ws, wi = torch_path.topk_oracle(scores, k)
        ok = torch.equal(mi, wi) and torch.equal(ms, ws)
        torch.save({"ok": ok, "ids": mi}, os.path.join(tmp, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gather_and_merge_world2_gloo(ured, tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert torch.equal(res[0]["ids"], res[1]["ids"])  # every rank ends with the same ranking
