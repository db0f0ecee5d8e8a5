"""Retrieval scoring and ranking on top of the Chamfer/DCD kernels.

Workloads: rank a partial target against K deformed candidates, or
against a whole source library, by cd_t ("cd_m" in the reference's pickles), cd_p ("cd_s") or
dcd -- what engine/generate_pair.py:69-122 does one B=1 call at a time through
engine/geometry_utils.py:80-82, followed by torch.topk(cd_m, k, largest=False)
(dataset/dataset_utils.py:1043-1051).

Orientation follows compute_dcd_loss(p1, p2) = calc_dcd(x=p1, gt=p2) with x = candidate /
library shape and gt = target: cloud 1 of the kernel is the target (broadcast over its
candidates, never copied), cloud 2 the candidate.

Sharded retrieval: the library is split contiguously over ranks (one process per GPU); every
rank scores its shard, takes a local top-k with global shape ids, and ONE all_gather of
[Q, k] (score, id) pairs over NCCL/NVLink merges them; all ranks end with the same ranking.
"""
import torch
import torch.distributed as dist

from . import _native
from .dist_chamfer_3D import _require_cloud, _stream


class PackedClouds:
    """[count, n, 3] float32 CUDA clouds plus their resident packed SoA image (the TMA source).

    The image is one independent block per cloud (include/ured_chamfer.h), so ``slice(lo, hi)`` is a
    zero-copy view: a library is packed once and scored slab by slab or shard by shard.
    """

    def __init__(self, xyz):
        if xyz.dim() == 2:
            xyz = xyz.unsqueeze(0)
        _require_cloud("xyz", xyz)
        self.xyz = xyz.contiguous()
        self.count, self.n, _ = self.xyz.shape
        lib = _native.load()
        nbytes = lib.ured_packed_bytes(self.count, self.n)
        self.packed = torch.empty(nbytes, device=self.xyz.device, dtype=torch.uint8)
        with torch.cuda.device(self.xyz.device):
            rc = lib.ured_pack_clouds(_native.ptr(self.xyz), self.count, self.n, None, _native.ptr(self.packed),
                                      _stream(self.xyz.device))
        _native.check(rc, "ured_pack_clouds")

    @property
    def device(self):
        return self.packed.device

    def repack_(self, xyz):
        """Pack new clouds of the same shape into this image's storage (in place, one kernel, on the current stream).

        After this the image no longer refers to a raw tensor (``self.xyz`` is None): the nearest-neighbour kernel reads
        queries and candidates from the image alone, so the caller's ``xyz`` may be freed or overwritten right away.
        Used by the retrieval engine: the image is a static buffer of its CUDA graph, the targets arrive in user tensors.
        """
        if tuple(xyz.shape) != (self.count, self.n, 3):
            raise ValueError(f"repack_ needs clouds of shape {(self.count, self.n, 3)}, got {tuple(xyz.shape)}")
        _require_cloud("xyz", xyz)
        xyz = xyz.contiguous()
        lib = _native.load()
        with torch.cuda.device(self.packed.device):
            rc = lib.ured_pack_clouds(_native.ptr(xyz), self.count, self.n, None, _native.ptr(self.packed), _stream(self.packed.device))
        _native.check(rc, "ured_pack_clouds")
        self.xyz = None
        return self

    @property
    def block_bytes(self):
        return (4 * ((self.n + 31) // 32 * 32) + 32) * 4

    def slice(self, lo, hi):
        """Clouds [lo, hi) as a PackedClouds sharing this one's storage."""
        if not (0 <= lo <= hi <= self.count):
            raise IndexError(f"slice [{lo}, {hi}) outside [0, {self.count})")
        view = object.__new__(PackedClouds)
        view.xyz = self.xyz[lo:hi] if self.xyz is not None else None
        view.count, view.n = hi - lo, self.n
        view.packed = self.packed[lo * self.block_bytes:]
        return view


def _as_packed(c):
    return c if isinstance(c, PackedClouds) else PackedClouds(c)


def nn_pairs(cloud1, cloud2, B, rep1, mod2, exact_only=False):
    """Nearest neighbours for B pairs: pair b = (cloud1[b // rep1], cloud2[b % mod2])."""
    lib = _native.load()
    c1, c2 = _as_packed(cloud1), _as_packed(cloud2)
    if (B + rep1 - 1) // rep1 > c1.count or min(B, mod2) > c2.count:
        raise ValueError("pair addressing runs past the clouds provided")
    dev = c1.device
    n1, n2 = c1.n, c2.n
    dist1 = torch.empty(B, n1, device=dev, dtype=torch.float32)
    dist2 = torch.empty(B, n2, device=dev, dtype=torch.float32)
    idx1 = torch.empty(B, n1, device=dev, dtype=torch.int32)
    idx2 = torch.empty(B, n2, device=dev, dtype=torch.int32)
    flags = _native.URED_FLAG_EXACT_ONLY if exact_only else 0
    scratch_bytes = lib.ured_nn_scratch_bytes(B, n1, n2)
    scratch = torch.empty(scratch_bytes, device=dev, dtype=torch.uint8) if scratch_bytes else None
    with torch.cuda.device(dev):
        rc = lib.ured_nn_packed(_native.ptr(c1.xyz), _native.ptr(c1.packed), n1,
                                _native.ptr(c2.xyz), _native.ptr(c2.packed), n2,
                                B, rep1, mod2, None, None,
                                _native.ptr(dist1), _native.ptr(dist2), _native.ptr(idx1), _native.ptr(idx2),
                                _native.ptr(scratch), scratch_bytes, flags, _stream(dev))
    _native.check(rc, "ured_nn_packed")
    return dist1, dist2, idx1, idx2


_METRICS = ("dcd", "cd_p", "cd_t")


def pair_scores(dist1, dist2, idx1, idx2, alpha=1000, n_lambda=1, exact_ranking=False, metrics=None):
    """(dcd, cd_p, cd_t), each [B], from raw NN results of chamfer(gt, x) (model_utils.py:31-58).

    Default: the fused epilogue kernel, whose row means follow torch's reduction order (include/ured_chamfer.h,
    "Reduction order") -- bit-identical to the reference's torch ops for the cloud sizes U-RED uses.
    ``exact_ranking=True`` runs the reference's torch ops themselves on the bit-exact dist/idx (model_utils.py:26-45,
    57-58: exp, scatter_add_, gather, pow, mean): ~25 small launches per call, for audits and for shapes outside the
    fused kernel's guaranteed range.

    ``metrics`` (subset of "dcd", "cd_p", "cd_t"; default all): scores not asked for come back as None -- ranking by
    cd_t alone skips the histograms and the exp/weight terms of the DCD loss."""
    metrics = _METRICS if metrics is None else tuple(metrics)
    if exact_ranking:
        from .model_utils import torch_epilogue
        n1, n2 = dist1.shape[1], dist2.shape[1]
        full = torch_epilogue(dist1, dist2, idx1, idx2, alpha, n_lambda, frac_12=n2 / n1, frac_21=n1 / n2)
        return tuple(v if m in metrics else None for m, v in zip(_METRICS, full))
    lib = _native.load()
    B, n1 = dist1.shape
    n2 = dist2.shape[1]
    dev = dist1.device
    out = torch.empty(3, B, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = lib.ured_dcd_forward(_native.ptr(dist1), _native.ptr(dist2), _native.ptr(idx1), _native.ptr(idx2),
                                  B, n1, n2, 1, max(B, 1), None, None,
                                  float(alpha), float(n_lambda), float(n2 / n1), float(n1 / n2), 0,
                                  *[_native.ptr(out[i]) if m in metrics else None for i, m in enumerate(_METRICS)],
                                  None, None, _stream(dev))
    _native.check(rc, "ured_dcd_forward")
    return tuple(out[i] if m in metrics else None for i, m in enumerate(_METRICS))


def score_candidates(targets, candidates, alpha=1000, n_lambda=1, exact_only=False, exact_ranking=False):
    """Score Q targets [Q, N, 3] against their K candidates [Q, K, M, 3].

    Returns {"dcd", "cd_p", "cd_t"}: each [Q, K], equal to calc_dcd(candidates[q, k], targets[q]).
    """
    if candidates.dim() != 4:
        raise ValueError("candidates must be [Q, K, M, 3]")
    Q, K, M, _ = candidates.shape
    if targets.shape[0] != Q:
        raise ValueError("one target per candidate set")
    cands = PackedClouds(candidates.reshape(Q * K, M, 3).float())
    tgts = _as_packed(targets.float() if not isinstance(targets, PackedClouds) else targets)
    raw = nn_pairs(tgts, cands, Q * K, K, Q * K, exact_only=exact_only)
    dcd, cd_p, cd_t = pair_scores(*raw, alpha=alpha, n_lambda=n_lambda, exact_ranking=exact_ranking)
    return {"dcd": dcd.view(Q, K), "cd_p": cd_p.view(Q, K), "cd_t": cd_t.view(Q, K)}


def score_library(targets, library, alpha=1000, n_lambda=1, max_pairs=8192, exact_only=False, exact_ranking=False, metrics=None):
    """Score every target [Q, N, 3] against every library shape: {"dcd","cd_p","cd_t"} each [Q, S] (or the subset ``metrics``).

    ``library`` is a PackedClouds (packed once, kept resident) or a [S, M, 3] tensor.  Work is cut
    into slabs of at most ``max_pairs`` (target, shape) pairs so that the per-point NN results
    (8 bytes per point per pair) stay a bounded scratch buffer instead of Q*S*(N+M)*8 bytes.
    """
    metrics = _METRICS if metrics is None else tuple(metrics)
    tgts = _as_packed(targets.float() if not isinstance(targets, PackedClouds) else targets)
    lib_c = _as_packed(library)
    Q, S = tgts.count, lib_c.count
    out = torch.empty(len(metrics), Q, S, device=tgts.device, dtype=torch.float32)
    result = {m: out[i] for i, m in enumerate(metrics)}
    if S == 0 or Q == 0:
        return result
    kw = dict(alpha=alpha, n_lambda=n_lambda, exact_ranking=exact_ranking, metrics=metrics)
    q_step = max(1, min(Q, max_pairs // S)) if S <= max_pairs else 1
    for q0 in range(0, Q, q_step):
        q1 = min(Q, q0 + q_step)
        sub_t = tgts.slice(q0, q1)
        if S <= max_pairs:
            raw = nn_pairs(sub_t, lib_c, (q1 - q0) * S, S, S, exact_only=exact_only)
            sc = dict(zip(_METRICS, pair_scores(*raw, **kw)))
            if q_step == Q:                      # one slab covers everything: hand the kernel's output back without a copy
                return {m: sc[m].view(Q, S) for m in metrics}
            for m in metrics:
                result[m][q0:q1] = sc[m].view(q1 - q0, S)
        else:  # one target at a time against slabs of the library
            for s0 in range(0, S, max_pairs):
                s1 = min(S, s0 + max_pairs)
                slab = lib_c.slice(s0, s1)
                raw = nn_pairs(sub_t, slab, s1 - s0, s1 - s0, s1 - s0, exact_only=exact_only)
                sc = dict(zip(_METRICS, pair_scores(*raw, **kw)))
                for m in metrics:
                    result[m][q0, s0:s1] = sc[m]
    return result


def score_all_pairs(library, alpha=1000, n_lambda=1, max_pairs=8192, exact_only=False):
    """All-pairs library scores, the batched form of `get_src_pair` (engine/generate_pair.py:69-85).

    Returns a [3, S, S] tensor (dcd, cd_p, cd_t) with entry [:, idx, i] = calc_dcd(x=library[i], gt=library[idx]) for
    i >= idx -- the row a reference pickle holds for shape idx -- and 0 below the diagonal (not evaluated).
    """
    lib_c = _as_packed(library)
    S = lib_c.count
    out = torch.zeros(3, S, S, device=lib_c.device, dtype=torch.float32)
    rows = max(1, max_pairs // max(S, 1))
    for q0 in range(0, S, rows):
        q1 = min(S, q0 + rows)
        sc = score_library(lib_c.slice(q0, q1), lib_c.slice(q0, S), alpha=alpha, n_lambda=n_lambda,
                           max_pairs=max_pairs, exact_only=exact_only)
        for m, key in enumerate(("dcd", "cd_p", "cd_t")):
            out[m, q0:q1, q0:] = sc[key]
    keep = torch.ones(S, S, device=out.device, dtype=torch.bool).triu()
    return out * keep


def write_pair_pickles(out_dir, names, scores):
    """Write one `<name>.pickle` per shape in the reference's layout (engine/generate_pair.py:82-85):
    {'dcd_loss', 'cd_s', 'cd_m'} -> float64 arrays over shapes idx..S-1, readable by read_pickle_topk
    (dataset/dataset_utils.py:1043-1051)."""
    import os
    import pickle
    import numpy as np
    sc = scores.detach().cpu().numpy().astype(np.float64)
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for idx, name in enumerate(names):
        rec = {"dcd_loss": sc[0, idx, idx:].copy(), "cd_s": sc[1, idx, idx:].copy(), "cd_m": sc[2, idx, idx:].copy()}
        path = os.path.join(out_dir, f"{name}.pickle")
        with open(path, "wb") as f:
            pickle.dump(rec, f)
        paths.append(path)
    return paths


def topk_smallest(scores, k, idx_offset=0):
    """k smallest per row in ascending (score, index) order -> (values [rows,k], indices int32 [rows,k]).

    torch.topk(scores, k, largest=False) with the tie order pinned to lowest index first.
    """
    if scores.dim() == 1:
        v, i = topk_smallest(scores.unsqueeze(0), k, idx_offset)
        return v[0], i[0]
    if not scores.is_cuda:
        raise RuntimeError("topk_smallest: GPU tensors only")
    lib = _native.load()
    scores = scores.contiguous().float()
    rows, cols = scores.shape
    vals = torch.empty(rows, k, device=scores.device, dtype=torch.float32)
    idx = torch.empty(rows, k, device=scores.device, dtype=torch.int32)
    with torch.cuda.device(scores.device):
        rc = lib.ured_topk_smallest(_native.ptr(scores), rows, cols, k, int(idx_offset),
                                    _native.ptr(vals), _native.ptr(idx), _stream(scores.device))
    _native.check(rc, "ured_topk_smallest")
    return vals, idx


def retrieve(targets, library, k=10, metric="cd_t", **score_kw):
    """Top-k library shapes per target: (scores [Q,k], shape ids int32 [Q,k])."""
    scores = score_library(targets, library, metrics=(metric,), **score_kw)[metric]
    return topk_smallest(scores, min(k, scores.shape[1]))


# ---- sharding over ranks ---------------------------------------------------------------------

def shard_bounds(num_shapes, world_size, rank):
    """Contiguous shard [lo, hi) of rank `rank`: ceil(S / world) shapes each, the tail may be short or empty."""
    per = (num_shapes + world_size - 1) // world_size
    lo = min(num_shapes, rank * per)
    return lo, min(num_shapes, lo + per)


def merge_topk(scores, ids, k):
    """Merge candidate lists [Q, C] by ascending (score, id); ids < 0 mark padding (native kernel, GPU tensors only)."""
    if not scores.is_cuda:
        raise RuntimeError("merge_topk: GPU tensors only")
    lib = _native.load()
    scores = scores.contiguous().float()
    ids = ids.contiguous().to(torch.int32)
    Q, C = scores.shape
    out_s = torch.empty(Q, k, device=scores.device, dtype=torch.float32)
    out_i = torch.empty(Q, k, device=scores.device, dtype=torch.int32)
    with torch.cuda.device(scores.device):
        rc = lib.ured_merge_topk(_native.ptr(scores), _native.ptr(ids), Q, C, k, _native.ptr(out_s), _native.ptr(out_i),
                                 _stream(scores.device))
    _native.check(rc, "ured_merge_topk")
    return out_s, out_i


def gather_and_merge(local_scores, local_ids, k, group=None, merge=None):
    """all_gather every rank's local top-k [Q, k_local<=k] (padded to k) and merge; same result on all ranks.

    This is the NCCL form of the exchange (three launches and one collective); `PeerExchange.topk` is the fused one.
    `merge` (default: the native merge_topk) exists so that the gloo-backed CPU tests can check the gather/padding
    logic with their own reference merge.
    """
    merge = merge or merge_topk
    Q = local_scores.shape[0]
    pad = k - local_scores.shape[1]
    if pad > 0:
        local_scores = torch.cat([local_scores, local_scores.new_full((Q, pad), float("inf"))], 1)
        local_ids = torch.cat([local_ids, local_ids.new_full((Q, pad), -1)], 1)
    # one message: (score bits, id) as int32 pairs
    msg = torch.stack([local_scores.contiguous().view(torch.int32), local_ids.to(torch.int32)], dim=-1).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        out = torch.empty((world * Q,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        dist.all_gather_into_tensor(out, msg, group=group)  # rank-major concatenation along dim 0
        out = out.view((world,) + tuple(msg.shape))
    else:
        out = msg.unsqueeze(0)
    allmsg = out.permute(1, 0, 2, 3).reshape(Q, -1, 2)
    return merge(allmsg[..., 0].contiguous().view(torch.float32), allmsg[..., 1].contiguous(), k)


def retrieve_sharded(targets, local_library, shard_offset, k=10, metric="cd_t", group=None, **score_kw):
    """Sharded top-k: `local_library` holds this rank's shapes, whose global ids start at `shard_offset`."""
    lib_c = _as_packed(local_library) if local_library is not None else None
    Q = targets.count if isinstance(targets, PackedClouds) else targets.shape[0]
    dev = targets.device
    if lib_c is None or lib_c.count == 0:
        ls = torch.empty(Q, 0, device=dev, dtype=torch.float32)
        li = torch.empty(Q, 0, device=dev, dtype=torch.int32)
    else:
        scores = score_library(targets, lib_c, metrics=(metric,), **score_kw)[metric]
        ls, li = topk_smallest(scores, min(k, lib_c.count), idx_offset=shard_offset)
    return gather_and_merge(ls, li, k, group=group)


class PendingResult:
    """Result of `RetrievalEngine.submit`: produced on one of the engine's lane streams; `result()` makes the caller's current
    stream wait for it (no host synchronisation) and hands the tensors over."""

    def __init__(self, tensors, event=None, stream=None):
        self._tensors, self._event, self._stream = tensors, event, stream

    def result(self):
        if self._event is not None:
            cur = torch.cuda.current_stream(self._tensors[0].device)
            cur.wait_event(self._event)
            for t in self._tensors:
                t.record_stream(cur)          # allocated on the lane stream, consumed on the caller's
            self._event = None
        return self._tensors


class RetrievalEngine:
    """Resident library shard + a CUDA graph of the whole per-query-batch pipeline.

    Retrieval against a sharded library is latency-bound when the shard is small (1000 shapes over 8 GPUs is
    ~0.15 ms of kernel time per query).  The engine keeps every buffer static and captures
        copy targets -> pack -> nn_kernel (+ split merge) -> dcd_fwd_kernel -> top-k / exchange
    once; ``query()`` copies the targets into the static input and replays the graph.

    exchange (only with more than one rank):
      "peer"  (default) one fused kernel: local top-k, NVLink stores of the [Q, k] keys into every peer's exchange buffer,
              flag wait, merge (`PeerExchange`, include/ured_chamfer.h "sharded retrieval") -- captured in the graph;
      "nccl"  top-k kernel, `all_gather_into_tensor`, merge kernel (the plain collective; also captured).
    Every rank ends with the same (scores [Q, k], global shape ids int32 [Q, k]).  Results are fresh tensors (the graph's
    static outputs are cloned), so a caller may hold them across queries.

    pipeline_depth > 1: that many lanes -- each its own stream, static buffers, graph and exchange buffers, all sharing the
    resident library -- take the submitted query batches in turn, so the next batch's nn_kernel fills the SMs while the previous
    batch's tail (the last partial wave of CTAs, dcd_fwd, the exchange's wait for its peers, the result copy) drains.  Use
    ``submit(targets) -> PendingResult`` to keep batches in flight; ``query()`` is submit + result (one batch at a time).
    Every rank must submit the same sequence of batches (the lanes exchange with their counterparts on the peers).
    """

    def __init__(self, local_library, shard_offset, num_queries, k=10, metric="cd_t", alpha=1000, n_lambda=1,
                 group=None, use_graph=True, max_pairs=16384, exchange="peer", exact_ranking=False, pipeline_depth=1):
        if local_library is None:
            self.lib = None
        elif isinstance(local_library, PackedClouds):
            self.lib = local_library if local_library.count else None
        else:
            self.lib = PackedClouds(local_library) if len(local_library) else None  # an empty shard holds nothing
        self.offset, self.Q, self.k, self.metric = int(shard_offset), int(num_queries), int(k), metric
        self.alpha, self.n_lambda, self.group, self.max_pairs = alpha, n_lambda, group, max_pairs
        self.exact_ranking = bool(exact_ranking)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange if self.world > 1 else "none"
        self.xchg = None
        self.graph = None
        self.static_in = None
        self.kernels_per_replay = 0
        self.use_graph = use_graph
        self.depth = max(1, int(pipeline_depth))
        self.lanes, self.lane_streams, self._next = [], [], 0
        if self.depth > 1:
            self.lanes = [RetrievalEngine(self.lib, shard_offset, num_queries, k=k, metric=metric, alpha=alpha, n_lambda=n_lambda, group=group,
                                          use_graph=use_graph, max_pairs=max_pairs, exchange=exchange, exact_ranking=exact_ranking)
                          for _ in range(self.depth)]

    def _peer(self, device):
        if self.xchg is None:
            from .exchange import PeerExchange
            self.xchg = PeerExchange(self.Q, self.k, device, group=self.group)
        return self.xchg

    def _scores(self, targets):
        if self.lib is None or self.lib.count == 0:
            return torch.empty(self.Q, 0, device=targets.device, dtype=torch.float32)   # (targets: a tensor or a PackedClouds)
        return score_library(targets, self.lib, alpha=self.alpha, n_lambda=self.n_lambda, max_pairs=self.max_pairs,
                             exact_ranking=self.exact_ranking, metrics=(self.metric,))[self.metric]

    def _pipeline(self, targets):
        Q, k = self.Q, self.k
        scores = self._scores(targets)
        if self.exchange == "peer":
            return self._peer(targets.device).topk(scores, self.offset)
        S_local = scores.shape[1]
        if S_local == 0:
            ls = torch.full((Q, k), float("inf"), device=targets.device)
            li = torch.full((Q, k), -1, device=targets.device, dtype=torch.int32)
        else:
            kk = min(k, S_local)
            ls, li = topk_smallest(scores, kk, idx_offset=self.offset)
            if kk < k:
                ls = torch.cat([ls, ls.new_full((Q, k - kk), float("inf"))], 1)
                li = torch.cat([li, li.new_full((Q, k - kk), -1)], 1)
        if self.exchange == "none":
            return ls, li
        msg = torch.stack([ls.contiguous().view(torch.int32), li], dim=-1).contiguous()
        out = torch.empty((self.world * Q, k, 2), dtype=torch.int32, device=msg.device)
        dist.all_gather_into_tensor(out, msg, group=self.group)
        allmsg = out.view(self.world, Q, k, 2).permute(1, 0, 2, 3).reshape(Q, -1, 2)
        return merge_topk(allmsg[..., 0].contiguous().view(torch.float32), allmsg[..., 1].contiguous(), k)

    def submit(self, targets, host_out=None):
        """Enqueue one query batch; returns a PendingResult.  With pipeline_depth lanes, consecutive batches overlap on the GPU.
        host_out: optional (scores, ids) pinned host tensors [Q, k]; the results are also copied there on the lane's stream."""
        if self.depth == 1:
            out = self._run(targets)
            if host_out is not None:
                host_out[0].copy_(out[0], non_blocking=True)
                host_out[1].copy_(out[1], non_blocking=True)
            return PendingResult(out)
        dev = targets.device
        if not self.lane_streams:
            self.lane_streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
        lane, stream = self.lanes[self._next], self.lane_streams[self._next]
        self._next = (self._next + 1) % self.depth
        stream.wait_stream(torch.cuda.current_stream(dev))      # the targets are ready on the caller's stream
        with torch.cuda.stream(stream):
            out = lane._run(targets)
            if host_out is not None:
                host_out[0].copy_(out[0], non_blocking=True)
                host_out[1].copy_(out[1], non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)
        targets.record_stream(stream)
        self.kernels_per_replay = lane.kernels_per_replay
        return PendingResult(out, done, stream)

    def drain(self):
        """Make the caller's current stream wait for everything submitted so far (no host synchronisation)."""
        for st in self.lane_streams:
            torch.cuda.current_stream(st.device).wait_stream(st)

    def query(self, targets):
        """targets [Q, N, 3] float32 CUDA -> (scores [Q, k], global shape ids int32 [Q, k]), identical on all ranks."""
        return self.submit(targets).result()

    def _run(self, targets):
        if targets.shape[0] != self.Q:
            raise ValueError(f"engine was built for {self.Q} queries per call")
        if not self.use_graph:
            return self._pipeline(targets.float())
        targets = targets.float()
        if self.graph is None:
            # the graph's static input is the targets' PACKED image: every query packs the caller's tensor straight into it
            # (one kernel, no staging copy) and the captured kernels read queries and candidates from the image alone
            self.static_in = PackedClouds(targets.contiguous())
            self.static_in.xyz = None
            if self.exchange == "peer":
                self._peer(targets.device)          # buffer mapping (collective set-up) happens outside the capture
            side = torch.cuda.Stream(device=targets.device)
            side.wait_stream(torch.cuda.current_stream(targets.device))
            with torch.cuda.stream(side):       # warm-up outside capture: attribute calls, NCCL connections, allocator pools
                for _ in range(3):
                    self._pipeline(self.static_in)
            torch.cuda.current_stream(targets.device).wait_stream(side)
            torch.cuda.synchronize(targets.device)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _native.load().ured_kernel_launches()
            with torch.cuda.graph(self.graph):
                self.static_out = self._pipeline(self.static_in)
            self.kernels_per_replay = int(_native.load().ured_kernel_launches() - n0) + 1  # library kernels per query (graph + the pack)
        self.static_in.repack_(targets)
        self.graph.replay()
        base = self.static_out[0]._base
        if base is not None and base is self.static_out[1]._base:   # both outputs live in one buffer: one copy instead of two
            both = base.clone()
            return both[0].view(torch.float32), both[1]
        return tuple(t.clone() for t in self.static_out)

    @property
    def exchange_mapping(self):
        x = self.lanes[0].xchg if self.lanes else self.xchg
        return x.mapping if x is not None else None

    def check(self):
        """Synchronise; raise if a peer exchange timed out (ids would be -2)."""
        for lane in self.lanes:
            lane.check()
        if self.xchg is not None:
            self.xchg.check()

    def close(self):
        for lane in self.lanes:
            lane.close()
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None
