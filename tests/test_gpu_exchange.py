"""GPU: the fused local top-k + peer exchange + merge kernel (ured_topk_exchange) and the sharded RetrievalEngine on it.

Three levels: (1) world = 1 through the raw C ABI against torch.sort; (2) several emulated ranks on ONE GPU -- every
"rank" is a buffer + a stream in this process, the kernels run concurrently and exchange through plain device pointers,
which exercises the whole flag / parity-slot / merge protocol without a second GPU; (3) real NCCL ranks with CUDA-IPC
mapped buffers when the box has at least two GPUs.
"""
import ctypes
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _alloc(lib, ured, world, rows, k):
    nbytes = lib.ured_xchg_bytes(world, rows, k)
    assert nbytes > 0
    p = ctypes.c_void_p()
    ured._native.check(lib.ured_xchg_alloc(nbytes, ctypes.byref(p)), "ured_xchg_alloc")
    return p.value


def _want(scores, k, offset=0):
    """ascending (score, index): torch.sort(stable=True); NaN last like the kernel's key order."""
    s, i = torch.sort(scores, dim=1, stable=True)
    return s[:, :k], (i[:, :k] + offset).int()


def test_exchange_world1_matches_sort(ured):
    lib = ured._native.load()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator().manual_seed(5)
    for rows, cols, k in [(1, 1000, 10), (64, 125, 10), (7, 37, 10), (3, 5, 10), (2, 0, 4), (5, 3000, 64)]:
        buf = _alloc(lib, ured, 1, rows, k)
        table = (ctypes.c_void_p * 1)(ctypes.c_void_p(buf))
        for rep in range(3):                                      # both parity slots and the epoch advance
            scores = torch.rand(rows, cols, generator=g).round(decimals=2).to(dev)   # heavy ties
            out_s = torch.empty(rows, k, device=dev)
            out_i = torch.empty(rows, k, device=dev, dtype=torch.int32)
            rc = lib.ured_topk_exchange(scores.data_ptr() if cols else None, rows, cols, k, 100, table, 1, 0, rows,
                                        out_s.data_ptr(), out_i.data_ptr(), 2000, stream)
            ured._native.check(rc, "ured_topk_exchange")
            torch.cuda.synchronize()
            kk = min(k, cols)
            ws, wi = _want(scores.cpu(), kk, 100)
            assert torch.equal(out_i[:, :kk].cpu(), wi) and torch.equal(out_s[:, :kk].cpu(), ws)
            assert (out_i[:, kk:] == -1).all() and torch.isinf(out_s[:, kk:]).all()
        status, epoch = ctypes.c_int(), ctypes.c_uint()
        ured._native.check(lib.ured_xchg_status(ctypes.c_void_p(buf), ctypes.byref(status), ctypes.byref(epoch), stream), "status")
        assert status.value == 0 and epoch.value == 3
        lib.ured_xchg_free(ctypes.c_void_p(buf))


def test_exchange_argument_errors(ured):
    lib = ured._native.load()
    assert lib.ured_xchg_bytes(17, 1, 10) == 0 and lib.ured_xchg_bytes(8, 513, 10) == 0 and lib.ured_xchg_bytes(8, 1, 65) == 0
    buf = _alloc(lib, ured, 2, 4, 10)
    table = (ctypes.c_void_p * 2)(ctypes.c_void_p(buf), None)
    out = torch.empty(4, 10, device="cuda")
    outi = torch.empty(4, 10, device="cuda", dtype=torch.int32)
    sc = torch.rand(4, 20, device="cuda")
    rc = lib.ured_topk_exchange(sc.data_ptr(), 4, 20, 10, 0, table, 2, 0, 4, out.data_ptr(), outi.data_ptr(), 100, None)
    assert rc == -1                                                # NULL peer buffer
    rc = lib.ured_topk_exchange(sc.data_ptr(), 4, 20, 10, 0, table, 2, 0, 8, out.data_ptr(), outi.data_ptr(), 100, None)
    assert rc == -4                                                # rows must match the buffer
    lib.ured_xchg_free(ctypes.c_void_p(buf))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_emulated_ranks_one_gpu(ured, world):
    """`world` ranks = `world` buffers + streams in this process: concurrent kernels exchanging through device pointers."""
    lib = ured._native.load()
    dev = torch.device("cuda", 0)
    rows, k, S = 16, 10, 1000
    g = torch.Generator().manual_seed(11)
    bufs = [_alloc(lib, ured, world, rows, k) for _ in range(world)]
    table = (ctypes.c_void_p * world)(*[ctypes.c_void_p(b) for b in bufs])
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    for rep in range(5):
        scores = torch.rand(rows, S, generator=g).round(decimals=3).to(dev)      # ties across shards
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            lo, hi = ured.shard_bounds(S, world, r)
            part = scores[:, lo:hi].contiguous()
            out_s = torch.empty(rows, k, device=dev)
            out_i = torch.empty(rows, k, device=dev, dtype=torch.int32)
            with torch.cuda.stream(streams[r]):
                rc = lib.ured_topk_exchange(part.data_ptr(), rows, hi - lo, k, lo, table, world, r, rows,
                                            out_s.data_ptr(), out_i.data_ptr(), 3000, streams[r].cuda_stream)
            ured._native.check(rc, "ured_topk_exchange")
            outs.append((out_s, out_i, part))
        torch.cuda.synchronize()
        ws, wi = _want(scores.cpu(), k)
        for out_s, out_i, _ in outs:
            assert torch.equal(out_i.cpu(), wi), "a rank's merged ids differ from the whole-library ranking"
            assert torch.equal(out_s.cpu(), ws)
    for b in bufs:
        status, epoch = ctypes.c_int(), ctypes.c_uint()
        ured._native.check(lib.ured_xchg_status(ctypes.c_void_p(b), ctypes.byref(status), ctypes.byref(epoch), None), "status")
        assert status.value == 0 and epoch.value == 5
        lib.ured_xchg_free(ctypes.c_void_p(b))


def test_exchange_missing_peer_times_out_instead_of_hanging(ured):
    lib = ured._native.load()
    dev = torch.device("cuda", 0)
    bufs = [_alloc(lib, ured, 2, 2, 4) for _ in range(2)]
    table = (ctypes.c_void_p * 2)(*[ctypes.c_void_p(b) for b in bufs])
    sc = torch.rand(2, 9, device=dev)
    out_s = torch.empty(2, 4, device=dev)
    out_i = torch.empty(2, 4, device=dev, dtype=torch.int32)
    rc = lib.ured_topk_exchange(sc.data_ptr(), 2, 9, 4, 0, table, 2, 0, 2, out_s.data_ptr(), out_i.data_ptr(), 200, None)
    ured._native.check(rc, "ured_topk_exchange")                   # rank 1 never calls
    torch.cuda.synchronize()
    assert (out_i == -2).all()
    status = ctypes.c_int()
    ured._native.check(lib.ured_xchg_status(ctypes.c_void_p(bufs[0]), ctypes.byref(status), None, None), "status")
    assert status.value == 1
    for b in bufs:
        lib.ured_xchg_free(ctypes.c_void_p(b))


# ---- real ranks ---------------------------------------------------------------------------------------------------
def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ured_b200 as ured
    from bench import library_rows, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        S, Q, n, k = 300, 3, 512, 10
        lo, hi = ured.shard_bounds(S, world, rank)
        shard = library_rows(lo, hi, n, dev)
        res = {}
        for mode, exchange, graph in [("peer-eager", "peer", False), ("peer-graph", "peer", True), ("nccl-eager", "nccl", False)]:
            eng = ured.RetrievalEngine(shard, lo, Q, k=k, metric="cd_t", use_graph=graph, exchange=exchange)
            ids = []
            for rep in range(4):
                _, tg = synth(Q, 8, n, seed=40 + rep)
                v, i = eng.query(tg.to(dev))
                ids.append(i.cpu())
            eng.check()
            res[mode] = torch.stack(ids)
            if exchange == "peer":
                res["mapping"] = eng.exchange_mapping
            eng.close()
        # two batches in flight: each lane exchanges through its own buffers with its counterpart on the peer
        eng = ured.RetrievalEngine(shard, lo, Q, k=k, metric="cd_t", exchange="peer", pipeline_depth=2)
        pend = []
        for rep in range(4):
            _, tg = synth(Q, 8, n, seed=40 + rep)
            pend.append(eng.submit(tg.to(dev)))
        eng.drain()
        res["peer-pipelined"] = torch.stack([p.result()[1].cpu() for p in pend])
        eng.check()
        eng.close()
        if rank == 0:
            full = library_rows(0, S, n, dev)
            want = []
            for rep in range(4):
                _, tg = synth(Q, 8, n, seed=40 + rep)
                want.append(ured.retrieve(tg.to(dev), full, k=k)[1].cpu())
            res["want"] = torch.stack(want)
        torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs on the box")
def test_sharded_engine_two_real_ranks(ured, tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 2000
    mp.start_processes(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    want = res[0]["want"]
    for r in res:
        for mode in ("peer-eager", "peer-graph", "nccl-eager", "peer-pipelined"):
            assert torch.equal(r[mode], want), f"{mode}: sharded ids differ from the one-rank ranking"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs on the box")
def test_large_shared_memory_kernels_on_two_devices_from_one_thread(ured):
    """The > 48 KB shared-memory opt-in is a per-device function attribute: one host thread calling on cuda:0 and then on
    cuda:1 must work on both (it used to be cached per thread, so the second device failed with cudaErrorInvalidValue)."""
    from conftest import make_clouds
    outs = []
    for d in (0, 1, 0):
        devd = torch.device("cuda", d)
        x = make_clouds(400, 2, 4096, "S").to(devd).requires_grad_()       # 4096 + 4096 points: 160 KB in the backward, 64 KB in dcd_fwd
        gt = make_clouds(401, 2, 4096, "S").to(devd).requires_grad_()
        loss, cd_p, cd_t = ured.calc_dcd(x, gt, alpha=100, n_lambda=1)
        loss.sum().backward()
        torch.cuda.synchronize(devd)
        outs.append((loss.detach().cpu(), x.grad.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])
