"""Host-side mirror of core/vectorindex.Hnsw (search side), over the C-ABI.

    vectorindex.Hnsw.Load     core/vectorindex/hnsw_commit.go:164-278  -> Hnsw.Load(blob)
    vectorindex.Hnsw.Search   core/vectorindex/hnsw.go:243-278         -> Hnsw.Search(query, k)
    vectorindex.SearchResult  core/vectorindex/hnsw_search_result.go   -> list of SearchResultItem

Insert/Remove (graph construction) stay with the reference for now; an index built and
Commit()ed by the Go side is loaded onto the GPU and searched there.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

from . import _lib

_u64p, _f32p, _i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int32)


@dataclass
class SearchResultItem:
    Id: int
    Score: float
    Metadata: dict = None


class Hnsw:
    def __init__(self, handle, dim=None):
        self._h = handle

    @staticmethod
    def Load(blob: bytes, device: int = 0) -> "Hnsw":
        """Hnsw.Load(r, header=true): a Commit(header=true) blob becomes a device-resident CSR graph."""
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
        h = C.c_void_p()
        _lib.check(_lib.lib().coltt_b200_hnsw_load(buf, len(blob), device, C.byref(h)))
        return Hnsw(h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().coltt_b200_hnsw_destroy(self._h)
            self._h = None

    __del__ = close

    def Len(self) -> int:
        n = C.c_uint64(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_len(self._h, C.byref(n)))
        return int(n.value)

    def BatchSearch(self, queries, k: int, ef: int = 0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        nq = q.shape[0]
        ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        _lib.check(_lib.lib().coltt_b200_hnsw_search(self._h, q.ctypes.data_as(_f32p), nq, int(k), int(ef), ids.ctypes.data_as(_u64p),
                                                      sc.ctypes.data_as(_f32p), cnt.ctypes.data_as(_i32p)))
        return ids, sc, cnt

    def Search(self, query, k: int, ef: int = 0) -> List[SearchResultItem]:
        """Hnsw.Search(ctx, query, k): nearest first, ascending Score (hnsw.go:268-275)."""
        ids, sc, cnt = self.BatchSearch(np.asarray(query, dtype=np.float32).reshape(1, -1), k, ef)
        return [SearchResultItem(int(ids[0, i]), float(sc[0, i])) for i in range(cnt[0])]

    def last_stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_last_stats(self._h, C.byref(a), C.byref(b)))
        return {"dist_evals": int(a.value), "expansions": int(b.value)}
