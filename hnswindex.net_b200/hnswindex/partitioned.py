"""
Partitioned index (mode 2): for graphs that exceed one GPU the data set is cut into
`world` contiguous id ranges, every rank (one process per GPU) builds / loads an independent
sub-HNSW over its range, all ranks see all queries, and the per-rank top-k lists are exchanged with
ONE all-gather (NCCL over NVLink on GPUs) and merged by `merge_keys_kernel` (K7) -- or, after `enable_fused()`, with
no collective at all: the search kernel's epilogue stores each query's top-k straight into every peer's gather buffer
(CUDA IPC peer pointers, NVLink stores overlapping the search of the next queries) and raises a per-rank flag; a wait
kernel + the merge kernel follow in stream order (csrc/partition_kernels.cu).

The reference has no distributed mode; parity for this mode is defined against the oracle running the
same sub-graphs + the same merge, and against brute force for recall (tests/test_partitioned_cpu.py,
tests/test_gpu_multi.py).

The host logic (range split, global ids, gather layout, merge order) is independent of the device
and is exercised on CPU with the gloo backend; `local_search` is pluggable for that purpose.
"""
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced id range [start, end) owned by `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _ordered(d: np.ndarray) -> np.ndarray:
    """Order-preserving uint32 image of fp32 (the kernels' key order: negative cosine distances sort first)."""
    b = np.ascontiguousarray(d, dtype=np.float32).view(np.uint32)
    return np.where(b >> 31 != 0, ~b, b | np.uint32(0x80000000)).astype(np.uint64)


def merge_topk_numpy(ids: np.ndarray, dists: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """CPU statement of merge_keys_kernel: ids/dists [parts, q, k'] -> [q, k], ascending by (distance, id),
    entries with id < 0 are padding and sort last; output padded with -1 / NaN."""
    parts, q, kk = ids.shape
    flat_ids = np.transpose(ids, (1, 0, 2)).reshape(q, parts * kk).astype(np.int64)
    flat_d = np.transpose(dists, (1, 0, 2)).reshape(q, parts * kk).astype(np.float32)
    key = (_ordered(flat_d) << np.uint64(32)) | (flat_ids.astype(np.uint64) & np.uint64(0x7FFFFFFF)) << np.uint64(1)
    key = np.where(flat_ids < 0, np.uint64(0xFFFFFFFFFFFFFFFF), key)
    order = np.argsort(key, axis=1, kind="stable")[:, :k]
    out_ids = np.take_along_axis(flat_ids, order, axis=1).astype(np.int32)
    out_d = np.take_along_axis(flat_d, order, axis=1)
    pad = out_ids < 0
    out_ids[pad] = -1
    out_d[pad] = np.nan
    if out_ids.shape[1] < k:
        extra = k - out_ids.shape[1]
        out_ids = np.concatenate([out_ids, np.full((q, extra), -1, np.int32)], axis=1)
        out_d = np.concatenate([out_d, np.full((q, extra), np.nan, np.float32)], axis=1)
    return out_ids, out_d


class PartitionedIndex:
    """One rank's view of a partitioned index.

    local_search(queries[nq, dim] float32, k) -> (local_ids[nq, k] int32, dists[nq, k] float32) searches this rank's
    sub-graph; ids are local slots (0-based inside the shard).  On a GPU rank pass an `hnswindex.Index`
    as `index` instead and the search, the id offset and the merge all stay on the device.
    """

    def __init__(self, rank: int, world: int, total: int, index=None, local_search: Optional[Callable] = None, group=None):
        if (index is None) == (local_search is None):
            raise ValueError("give exactly one of index / local_search")
        self.rank, self.world, self.total = rank, world, total
        self.start, self.end = shard_range(total, rank, world)
        self.index, self.local_search, self.group = index, local_search, group
        self._fused = None

    # ---- device-independent pieces ------------------------------------------------------------
    def to_global(self, local_ids: np.ndarray) -> np.ndarray:
        return np.where(local_ids >= 0, local_ids + self.start, -1).astype(np.int32)

    def _gather(self, ids_t, d_t):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return ids_t.unsqueeze(0), d_t.unsqueeze(0)
        all_ids = [torch.empty_like(ids_t) for _ in range(self.world)]
        all_d = [torch.empty_like(d_t) for _ in range(self.world)]
        dist.all_gather(all_ids, ids_t, group=self.group)
        dist.all_gather(all_d, d_t, group=self.group)
        return torch.stack(all_ids), torch.stack(all_d)

    # ---- query -----------------------------------------------------------------------------------
    def knn_query(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Every rank passes the SAME queries; every rank returns the same merged global top-k."""
        import torch
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if self.index is not None and torch.cuda.is_available():
            return self._knn_query_cuda(q, k)
        ids, d = self.local_search(q, k)
        ids_t = torch.from_numpy(self.to_global(np.asarray(ids)))
        d_t = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
        g_ids, g_d = self._gather(ids_t, d_t)
        return merge_topk_numpy(g_ids.numpy(), g_d.numpy(), k)

    # ---- fused search + exchange (no collective in the data path) ----------------------------------------
    def enable_fused(self, nq_max: int, k: int) -> None:
        """Allocate this rank's double-buffered gather buffer + flags, exchange the CUDA IPC handles with the other ranks
        (a one-off all_gather_object) and map every peer's buffer.  All ranks must call it with the same arguments."""
        import ctypes as ct
        import torch
        import torch.distributed as dist
        from . import bindings as hb
        w = self.world
        half = w * nq_max * k                       # elements of one parity of ids (and of dists)
        flags_off = 2 * half * 4 * 2                # bytes: ids[2][half] int32, dists[2][half] fp32, then flags[world] u32
        nbytes = flags_off + 256
        ptr = ct.c_void_p()
        handle = (ct.c_ubyte * 64)()
        hb._check(hb.lib.hnsw_ipc_alloc(ct.c_size_t(nbytes), ct.byref(ptr), handle))
        handles = [None] * w
        if w > 1:
            dist.all_gather_object(handles, bytes(handle), group=self.group)
        else:
            handles[0] = bytes(handle)
        base = []
        for p in range(w):
            if p == self.rank:
                base.append(ptr.value)
            else:
                pp = ct.c_void_p()
                hb._check(hb.lib.hnsw_ipc_open((ct.c_ubyte * 64).from_buffer_copy(handles[p]), ct.byref(pp)))
                base.append(pp.value)
        self._fused = dict(nq_max=nq_max, k=k, base=base, half=half, flags_off=flags_off, epoch=0, own=ptr.value,
                           out_ids=torch.empty((nq_max, k), dtype=torch.int32, device="cuda"),
                           out_d=torch.empty((nq_max, k), dtype=torch.float32, device="cuda"))
        if w > 1:
            dist.barrier(group=self.group)          # every buffer is mapped before anybody writes into it

    def disable_fused(self) -> None:
        """Unmap the peers' buffers and free this rank's (all ranks together: a barrier separates unmapping from freeing)."""
        import torch.distributed as dist
        from . import bindings as hb
        f, self._fused = self._fused, None
        if f is None:
            return
        for p, ptr in enumerate(f["base"]):
            if p != self.rank:
                hb.lib.hnsw_ipc_close(ptr)
        if self.world > 1:
            dist.barrier(group=self.group)
        hb.lib.hnsw_ipc_free(f["own"])

    def _knn_query_fused(self, q_dev, nq: int, k: int, sync: bool = True):
        import ctypes as ct
        from . import bindings as hb
        f = self._fused
        if f.get("broken"):
            raise RuntimeError("fused partitioned query: a peer timed out earlier and the epoch protocol is out of step; "
                               "call disable_fused() and enable_fused() on ALL ranks before the next fused query")
        if nq > f["nq_max"] or k != f["k"]:
            raise ValueError("fused partitioned query: batch larger than the buffers of enable_fused(), or a different k")
        f["epoch"] += 1
        par = f["epoch"] & 1
        sc = hb.Scatter()
        sc.world, sc.rank, sc.idOffset, sc.epoch = self.world, self.rank, self.start, f["epoch"]
        # layout of a buffer: ids[2][world][nq_max][k] | dists[2][world][nq_max][k] | flags[world]; a query batch of nq rows
        # uses the first nq rows of every rank's slot, so the slot pitch is nq (the kernel indexes [rank][nq][k])
        ids_off = par * f["half"] * 4
        d_off = 2 * f["half"] * 4 + par * f["half"] * 4
        for p in range(self.world):
            sc.ids[p] = ct.cast(ct.c_void_p(f["base"][p] + ids_off), ct.POINTER(ct.c_int))
            sc.dists[p] = ct.cast(ct.c_void_p(f["base"][p] + d_off), ct.POINTER(ct.c_float))
            sc.flags[p] = ct.cast(ct.c_void_p(f["base"][p] + f["flags_off"]), ct.POINTER(ct.c_uint))
        hb._check(hb.lib.hnsw_knn_query_scatter(self.index._h, ct.c_void_p(q_dev.data_ptr()), nq, self.index.dim, k, ct.byref(sc)))
        fn = hb.lib.hnsw_wait_and_merge_device if sync else hb.lib.hnsw_wait_and_merge_device_async
        rc = fn(self.index._h, ct.c_void_p(f["own"] + f["flags_off"]), self.world, f["epoch"],
                ct.c_void_p(f["own"] + ids_off), ct.c_void_p(f["own"] + d_off), nq, k,
                ct.c_void_p(f["out_ids"].data_ptr()), ct.c_void_p(f["out_d"].data_ptr()))
        if rc < 0:
            f["broken"] = True      # a peer never delivered: outputs are -1 / NaN, the exchange must be re-established
            hb._check(rc)
        return f["out_ids"][:nq], f["out_d"][:nq]

    def sync(self) -> None:
        """Wait for the batches queued with knn_query_device(..., sync=False); raises if a peer never delivered."""
        from . import bindings as hb
        rc = hb.lib.hnsw_partition_sync(self.index._h)
        if rc < 0 and self._fused is not None:
            self._fused["broken"] = True
        hb._check(rc)

    def set_peer_timeout_ms(self, ms: int) -> None:
        """How long the wait+merge kernel waits (wall time, on the device) for a peer GPU before it gives up: the batch's
        outputs become -1 / NaN, sync() raises, and the exchange has to be re-established.  Default 10 s."""
        self.index.set_tuning("peer_timeout_ms", int(ms))

    def knn_query_device(self, q_dev, k: int, sync: bool = True):
        """Queries already on this GPU (torch tensor [nq, dim] fp32) -> merged (ids, dists) torch tensors on the GPU.
        sync=False (fused path only): the batch is only queued on the index's stream -- call sync() before reading the
        result tensors, which the next batch overwrites."""
        nq = int(q_dev.shape[0])
        if self._fused is not None:
            return self._knn_query_fused(q_dev, nq, k, sync)
        return self._knn_query_gather(q_dev, nq, k)

    def _knn_query_cuda(self, q: np.ndarray, k: int):
        import torch
        ids, d = self.knn_query_device(torch.from_numpy(q).cuda(), k)
        return ids.cpu().numpy(), d.cpu().numpy()

    def _knn_query_gather(self, q_dev, nq: int, k: int):
        import ctypes as ct
        import torch
        from . import bindings as hb
        ids_dev = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        d_dev = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        self.index.knn_query_device(q_dev.data_ptr(), nq, k, ids_dev.data_ptr(), d_dev.data_ptr())
        ids_dev = torch.where(ids_dev >= 0, ids_dev + self.start, ids_dev)
        g_ids, g_d = self._gather(ids_dev.contiguous(), d_dev)
        out_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        out_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        torch.cuda.current_stream().synchronize()
        rc = hb.lib.hnsw_merge_topk_device(ct.c_void_p(g_ids.contiguous().data_ptr()), ct.c_void_p(g_d.contiguous().data_ptr()), self.world, nq, k,
                                           ct.c_void_p(out_ids.data_ptr()), ct.c_void_p(out_d.data_ptr()), None)
        if rc < 0:
            raise RuntimeError(hb.last_error())
        return out_ids, out_d
