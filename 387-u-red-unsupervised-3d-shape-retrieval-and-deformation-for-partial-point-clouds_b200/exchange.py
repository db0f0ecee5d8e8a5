"""Peer-memory exchange for sharded retrieval: local top-k + NVLink stores into every peer + merge, ONE kernel.

SURVEY.md 8(e): the library is split over the ranks of one box (one process per GPU), each rank ranks its shard, and
the global top-k needs one exchange of [Q, k] (score, id) pairs.  The plain way -- top-k kernel, NCCL all_gather, merge
kernel -- costs three launches and a collective's latency for an 80-byte message; `PeerExchange.topk` does the whole
step in one kernel over peer-mapped buffers (include/ured_chamfer.h, "sharded retrieval").  torch.distributed is used
once, at set-up, to hand the 64-byte buffer handles around and as a barrier.

The reference has no multi-GPU retrieval (engine/generate_pair.py:69-122 is a single-GPU loop followed by
torch.topk, dataset/dataset_utils.py:1043-1051); the result here is defined as that ranking over the whole library.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _native


class PeerExchange:
    """Exchange buffers of one process group, mapped into every rank.

    rows / k are fixed per instance (every rank must build it with the same values and call `topk` in lock step).
    Mapping: CUDA IPC handles of a buffer allocated by the library (`ured_xchg_alloc`); if the platform refuses IPC,
    torch symmetric memory provides the peer pointers instead.  Both fail -> NativeLibraryError (no silent fallback:
    the caller chooses exchange="nccl" explicitly if it wants the collective).
    """

    def __init__(self, rows, k, device, group=None, timeout_ms=20000):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.lib = _native.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.k, self.device = int(rows), int(k), torch.device(device)
        self.timeout_ms = int(timeout_ms)
        self.nbytes = self.lib.ured_xchg_bytes(self.world, self.rows, self.k)
        if self.nbytes == 0:
            raise ValueError(f"peer exchange supports world <= 16, rows <= 512, k <= 64 (got {self.world}, {rows}, {k})")
        self.own = None
        self.peers = [None] * self.world
        self._symm = None
        self.mapping = None
        with torch.cuda.device(self.device):
            ok = self._all_agree(self._map_ipc)
            if not ok:
                self._unmap_ipc()
                ok = self._all_agree(self._map_symm)
                if not ok:
                    raise _native.NativeLibraryError("peer exchange: neither CUDA IPC nor torch symmetric memory could map the buffers "
                                                     f"(last error: {self.lib.ured_last_error_string().decode()})")
        self._table = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(p) for p in self.peers])
        dist.barrier(group=self.group)   # every rank has mapped every buffer before the first store can arrive

    # ---- mapping ------------------------------------------------------------------------------------------------
    def _all_agree(self, fn):
        """Run fn() on every rank; the mapping counts only if it worked everywhere."""
        try:
            ok = bool(fn())
        except Exception:  # noqa: BLE001 - any failure means "try the next mapping", on every rank alike
            ok = False
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def _map_ipc(self):
        lib = self.lib
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        rc = lib.ured_xchg_alloc(self.nbytes, ctypes.byref(ptr))   # (a failing rank still takes part in the all_gather below)
        if rc == 0:
            self.own = ptr.value
            rc = lib.ured_xchg_export(ctypes.c_void_p(self.own), handle)
        mine = torch.tensor(list(handle) + [0 if rc == 0 else 1], dtype=torch.uint8, device=self.device)
        every = torch.empty(self.world * 65, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.cpu().view(self.world, 65)
        if int(every[:, 64].max()) != 0:
            return False
        self.peers = [None] * self.world
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self.peers[r] = self.own
                continue
            buf = (ctypes.c_ubyte * 64)(*every[r, :64].tolist())
            out = ctypes.c_void_p()
            if lib.ured_xchg_import(buf, ctypes.byref(out)) != 0:
                return False
            self.peers[r] = out.value
            self._opened.append(out.value)
        self.mapping = "cuda-ipc"
        return True

    def _unmap_ipc(self):
        for p in getattr(self, "_opened", []):
            self.lib.ured_xchg_close(ctypes.c_void_p(p))
        self._opened = []
        if self.own is not None and self._symm is None:
            self.lib.ured_xchg_free(ctypes.c_void_p(self.own))
        self.own = None

    def _map_symm(self):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        t.zero_()
        torch.cuda.synchronize(self.device)
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm.rendezvous(t, grp)
        self._symm = (t, hdl)
        self.peers = [int(p) for p in hdl.buffer_ptrs]
        self.own = self.peers[self.rank]
        self.mapping = "torch-symmetric-memory"
        return True

    # ---- the exchange -------------------------------------------------------------------------------------------
    def topk(self, scores, idx_offset):
        """scores [rows, S_local] float32 CUDA (S_local may be 0) -> (scores [rows, k], global ids int32 [rows, k]),
        the k smallest over all ranks in ascending (score, id) order; the same tensors on every rank."""
        if scores.dim() != 2 or scores.shape[0] != self.rows:
            raise ValueError(f"exchange was built for {self.rows} rows, got {tuple(scores.shape)}")
        if not scores.is_cuda:
            raise RuntimeError("PeerExchange.topk: GPU tensors only")
        scores = scores.contiguous().float()
        both = torch.empty(2, self.rows, self.k, device=scores.device, dtype=torch.int32)   # one buffer: [0] score bits, [1] ids
        out_s, out_i = both[0].view(torch.float32), both[1]
        with torch.cuda.device(scores.device):
            rc = self.lib.ured_topk_exchange(_native.ptr(scores) if scores.numel() else None, self.rows, scores.shape[1], self.k,
                                             int(idx_offset), self._table, self.world, self.rank, self.rows,
                                             _native.ptr(out_s), _native.ptr(out_i), self.timeout_ms,
                                             torch.cuda.current_stream(scores.device).cuda_stream)
        _native.check(rc, "ured_topk_exchange")
        return out_s, out_i

    def check(self):
        """Synchronise and raise if any exchange so far gave up waiting for a peer."""
        status, epoch = ctypes.c_int(), ctypes.c_uint()
        with torch.cuda.device(self.device):
            _native.check(self.lib.ured_xchg_status(ctypes.c_void_p(self.own), ctypes.byref(status), ctypes.byref(epoch),
                                                    torch.cuda.current_stream(self.device).cuda_stream), "ured_xchg_status")
        if status.value != 0:
            raise _native.NativeLibraryError(f"peer exchange timed out waiting for a rank (after {epoch.value} exchanges)")
        return epoch.value

    def close(self):
        if self.own is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)      # nobody is still storing into a buffer that is about to go away
        if self._symm is None:
            self._unmap_ipc()
        self._symm = None
        self.own = None
