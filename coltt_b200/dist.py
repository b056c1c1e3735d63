"""Sharded search across the GPUs of one node: one process per GPU, rows partitioned one shard per
GPU, ONE exchange step per batch — an all-gather of the per-shard top-k — then the K5 merge.

Reference analogue: the 16 in-process shards of a vectorspace, each scanned into a shard-local queue
and re-merged (edge/none_vectorstore.go:152-178); rows -> shard by pkg/sharding.ShardVertex
(pkg/sharding/shard.go:34-41).  There is no other collective in the path.

torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is plumbing only: the local
search and the merge are injected callables — on a GPU box they are the C-ABI calls
coltt_b200_store_search_dev / coltt_b200_merge_topk_dev.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)

# one hit on the wire == coltt_hit (include/coltt_b200.h): u64 id, f32 score, u32 slot
HIT_DTYPE = np.dtype([("id", "<u8"), ("score", "<f4"), ("slot", "<u4")])


def shard_vertex(ids, c: int = 16) -> np.ndarray:
    """pkg/sharding.ShardVertex vectorised: FNV-1a 64 over the little-endian bytes of the id, mod c."""
    x = np.ascontiguousarray(ids, dtype=np.uint64)
    h = np.full(x.shape, FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(8):
            h ^= (x >> np.uint64(8 * i)) & np.uint64(0xFF)
            h *= FNV_PRIME
    return (h % np.uint64(c)).astype(np.int64)


def gpu_of(ids, world: int) -> np.ndarray:
    """Row -> GPU: the reference's shard identity folded onto the GPUs (ShardVertex(id,16) mod G)."""
    return shard_vertex(ids, 16) % world


class ShardedSearch:
    """Per-rank driver of the sharded FLAT search.

    local_search(queries[nq,dim] f32, k, select_mode) -> (hits[nq,k] HIT_DTYPE-shaped tensor, counts[nq] int32 tensor)
    merge(gathered_hits[world,nq,k], gathered_counts[world,nq], k, select_mode) -> (hits[nq,k], counts[nq])
    Both work on torch tensors living where the process group communicates (CUDA for NCCL, CPU for gloo);
    hits travel as int32 [..., 4] views of coltt_hit.
    """

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local_search = local_search
        self.merge = merge

    def search(self, queries, k: int, select_mode: int):
        import torch
        hits, counts = self.local_search(queries, k, select_mode)
        if self.world == 1:
            return self.merge(hits.unsqueeze(0), counts.unsqueeze(0), k, select_mode)
        nq = hits.shape[0]
        # ONE message per shard: its [nq,k] hits followed by its counts (padded to a 16-byte multiple);
        # the gathered buffer is rank-major, and the merge sees strided views of it (no repacking)
        hflat = hits.reshape(-1)
        n_h, n_c = hflat.numel(), (nq + 3) // 4 * 4
        packed = torch.zeros((n_h + n_c,), dtype=hits.dtype, device=hits.device)
        packed[:n_h].copy_(hflat)
        packed[n_h:n_h + nq].copy_(counts.to(hits.dtype))
        gathered = torch.empty((self.world * (n_h + n_c),), dtype=hits.dtype, device=hits.device)
        # the one exchange step: per-shard top-k of every rank to every rank
        self.dist.all_gather_into_tensor(gathered, packed, group=self.group)
        g2 = gathered.view(self.world, n_h + n_c)
        g_hits = g2[:, :n_h].unflatten(1, tuple(hits.shape))
        g_counts = g2[:, n_h:n_h + nq]
        return self.merge(g_hits, g_counts, k, select_mode)


def cuda_callables(space, device_index: int, math_mode: Optional[int] = None, stream=None):
    """(local_search, merge) bound to a coltt_b200 VectorSpace on one GPU through the C-ABI."""
    import torch
    from . import _lib
    L = _lib.lib()
    dev = torch.device("cuda", device_index)
    mm = space.math_mode if math_mode is None else math_mode

    # The library must be handed a REAL stream: pointer 0 means "use your own private stream", which is not ordered with
    # torch's legacy default stream — the copies that produce `q` (and any fill of the outputs) could still be in flight when
    # the library's kernels start.  So the calls run on a side stream that is made to wait for torch's current stream, and
    # torch's current stream waits for it afterwards.
    side = stream or torch.cuda.Stream(device=dev)

    def _ordered(fn):
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        fn(side.cuda_stream)
        cur.wait_stream(side)

    def local_search(queries, k, select_mode):
        q = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(queries, np.float32))
        q = q.to(dev, dtype=torch.float32).contiguous()
        nq = q.shape[0]
        hits = torch.empty((nq, k, 4), dtype=torch.int32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        _ordered(lambda sp: _lib.check(L.coltt_b200_store_search_dev(space._h, q.data_ptr(), nq, k, select_mode, mm, hits.data_ptr(),
                                                                     counts.data_ptr(), sp)))
        q.record_stream(side)
        return hits, counts

    def merge(gathered, gcounts, k, select_mode):
        world, nq = gathered.shape[0], gathered.shape[1]
        out = torch.empty((nq, k, 4), dtype=torch.int32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        if gathered.is_contiguous() and gcounts.is_contiguous():
            _ordered(lambda sp: _lib.check(L.coltt_b200_merge_topk_dev(device_index, gathered.data_ptr(), gcounts.data_ptr(), world, nq,
                                                                       gathered.shape[2], k, select_mode, out.data_ptr(), cnt.data_ptr(), sp)))
        else:   # strided views of the packed all-gather buffer: one block per shard, hits then counts
            stride_b = gathered.stride(0) * 4
            off_b = gcounts.data_ptr() - gathered.data_ptr()
            assert gcounts.stride(0) * 4 == stride_b and off_b > 0
            _ordered(lambda sp: _lib.check(L.coltt_b200_merge_topk_dev2(device_index, gathered.data_ptr(), world, nq, gathered.shape[2], k,
                                                                        select_mode, stride_b, off_b, out.data_ptr(), cnt.data_ptr(), sp)))
        return out, cnt

    return local_search, merge


class Comm:
    """One rank of a sharded collection behind the C-ABI (include/coltt_b200.h, csrc/comm.cu): the local search, the single
    exchange of per-shard top-k (fused into the merge over NVLink peer memory, else one NCCL all-gather) and the K5 merge all
    run inside libcoltt_b200.so on the rank's stream, with persistent exchange buffers and pinned staging.  torch.distributed
    is used only to hand the 128-byte rendezvous blob to the ranks."""

    def __init__(self, handle, rank: int, world: int, device: int):
        self._h, self.rank, self.world, self.device = handle, rank, world, device

    @classmethod
    def from_torch_distributed(cls, device_index: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        blob = (C.c_uint8 * 128)()
        if rank == 0 and world > 1:
            _lib.check(L.coltt_b200_comm_unique_id(blob, 128))
        if world > 1:
            t = torch.tensor(list(blob), dtype=torch.uint8)
            t = t.cuda(device_index) if dist.get_backend(group) == "nccl" else t
            dist.broadcast(t, src=0, group=group)
            blob = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        if world > 1:
            _lib.check(L.coltt_b200_comm_init_rank(blob, rank, world, device_index, C.byref(h)))
        else:
            dev = (C.c_int * 1)(device_index)
            _lib.check(L.coltt_b200_init(dev, 1, C.byref(h)))
        return cls(h, rank, world, device_index)

    def close(self):
        if getattr(self, "_h", None):
            from . import _lib
            _lib.lib().coltt_b200_comm_destroy(self._h)
            self._h = None

    @property
    def exchange(self) -> str:
        """"peer-memory" | "nccl" | "undecided" (before the first sharded search): coltt_b200_comm_exchange_mode."""
        from . import _lib
        m = _lib.lib().coltt_b200_comm_exchange_mode(self._h)
        return {1: "peer-memory", 2: "nccl"}.get(m, "undecided")

    def search(self, space, queries, k: int, select_mode: int, math_mode: Optional[int] = None):
        """Collective: every rank calls with the same queries.  Host buffers in, merged (ids, scores, counts) out."""
        import ctypes as C
        from . import _lib
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq = q.shape[0]
        ids = np.zeros((nq, k), np.uint64)
        sc = np.zeros((nq, k), np.float32)
        cnt = np.zeros(nq, np.int32)
        mm = space.math_mode if math_mode is None else math_mode
        _lib.check(_lib.lib().coltt_b200_sharded_search(self._h, space._h, q.ctypes.data_as(C.POINTER(C.c_float)), nq, k, select_mode, mm,
                                                        ids.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)),
                                                        cnt.ctypes.data_as(C.POINTER(C.c_int32))))
        return ids, sc, cnt

    def search_dev(self, space, d_queries_ptr: int, nq: int, k: int, select_mode: int, math_mode: int, d_out_ptr: int, d_counts_ptr: int,
                   stream_ptr: int = 0):
        from . import _lib
        _lib.check(_lib.lib().coltt_b200_sharded_search_dev(self._h, space._h, d_queries_ptr, nq, k, select_mode, math_mode, d_out_ptr,
                                                            d_counts_ptr, stream_ptr))

    def hnsw_search(self, sub_graph, queries, k: int, ef: int = 0, pq: bool = False, rerank: bool = True):
        """Collective Hnsw.Search over one sub-graph per GPU; pq = the product-quantized walk (config 5)."""
        import ctypes as C
        from . import _lib
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq = q.shape[0]
        ids = np.zeros((nq, k), np.uint64)
        sc = np.zeros((nq, k), np.float32)
        cnt = np.zeros(nq, np.int32)
        L = _lib.lib()
        outs = (ids.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_int32)))
        if pq:
            _lib.check(L.coltt_b200_sharded_hnsw_pq_search(self._h, sub_graph._h, q.ctypes.data_as(C.POINTER(C.c_float)), nq, k, ef, 1 if rerank else 0, *outs))
        else:
            _lib.check(L.coltt_b200_sharded_hnsw_search(self._h, sub_graph._h, q.ctypes.data_as(C.POINTER(C.c_float)), nq, k, ef, *outs))
        return ids, sc, cnt


def unpack_hits(hits, counts):
    """int32 [nq,k,4] coltt_hit tensor -> (ids u64 [nq,k], scores f32 [nq,k], counts)."""
    h = hits.detach().cpu().contiguous().numpy().view(HIT_DTYPE).reshape(hits.shape[0], hits.shape[1])
    return h["id"].copy(), h["score"].copy(), counts.detach().cpu().numpy()
