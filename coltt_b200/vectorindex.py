"""Host-side mirror of core/vectorindex.Hnsw (search side), over the C-ABI.

    vectorindex.Hnsw.Load     core/vectorindex/hnsw_commit.go:164-278  -> Hnsw.Load(blob)
    vectorindex.Hnsw.Search   core/vectorindex/hnsw.go:243-278         -> Hnsw.Search(query, k)
    vectorindex.SearchResult  core/vectorindex/hnsw_search_result.go   -> list of SearchResultItem

    n x vectorindex.Hnsw.Insert  core/vectorindex/hnsw.go:104-167         -> Hnsw.Build(ids, vecs, ...) (bulk, on the GPU)
    vectorindex.Hnsw.Commit   core/vectorindex/hnsw_commit.go:69-162   -> Hnsw.Commit()

An index built and Commit()ed by the Go side is loaded onto the GPU and searched there; an initial load can
instead be built on the GPU (csrc/hnsw_build.cu) and handed back to the Go side as a Commit blob.  Incremental
Insert/Remove stay with the reference.
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

    @staticmethod
    def Build(ids, vecs, metric: int = 0, m: int = 16, ef: int = 20, ef_construction: int = 200, levels=None, seed: int = 0xC0177,
              device: int = 0) -> "Hnsw":
        """Bulk equivalent of NewHnsw(dim, space, WithM(m), ...) followed by Insert(id, vec, md, level) for every row
        (hnsw.go:56-83,104-167).  `levels` = the vertexLevel per row the caller would pass to Insert, or None to draw
        them as Hnsw.RandomLevel does (hnsw.go:280-282) from `seed`."""
        v = np.ascontiguousarray(vecs, dtype=np.float32)
        i = np.ascontiguousarray(ids, dtype=np.uint64)
        if v.ndim != 2 or v.shape[0] != i.shape[0]:
            raise ValueError("ids / vecs shape mismatch")
        cfg = _lib.HnswBuildCfg(v.shape[1], int(metric), int(m), int(ef), int(ef_construction), int(device), int(seed))
        lv = None
        if levels is not None:
            lv = np.ascontiguousarray(levels, dtype=np.int32)
            if lv.shape[0] != i.shape[0]:
                raise ValueError("levels shape mismatch")
        h = C.c_void_p()
        _lib.check(_lib.lib().coltt_b200_hnsw_build(C.byref(cfg), i.ctypes.data_as(_u64p), v.ctypes.data_as(_f32p),
                                                     lv.ctypes.data_as(_i32p) if lv is not None else None, v.shape[0], C.byref(h)))
        return Hnsw(h)

    def Commit(self) -> bytes:
        """Hnsw.Commit(w, header=true): the reference's index blob (loadable by the Go side and by Hnsw.Load)."""
        n = C.c_size_t(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_commit(self._h, None, C.byref(n)))
        buf = (C.c_uint8 * max(n.value, 1))()
        _lib.check(_lib.lib().coltt_b200_hnsw_commit(self._h, buf, C.byref(n)))
        return bytes(memoryview(buf)[: n.value])

    def build_stats(self):
        ms = (C.c_double * 4)()
        ne, ml = C.c_uint64(0), C.c_int32(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_build_stats(self._h, ms, C.byref(ne), C.byref(ml)))
        fs = (C.c_uint64 * 2)()
        _lib.check(_lib.lib().coltt_b200_hnsw_build_fast_stats(self._h, fs))
        return {"ingest_ms": ms[0], "knn_ms": ms[1], "edge_dist_ms": ms[2], "host_graph_ms": ms[3], "n_edges": int(ne.value),
                "max_level": int(ml.value), "fast_queries": int(fs[0]), "fast_fallbacks": int(fs[1])}

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().coltt_b200_hnsw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Len(self) -> int:
        n = C.c_uint64(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_len(self._h, C.byref(n)))
        return int(n.value)

    def Dim(self) -> int:
        d = C.c_uint32(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_dim(self._h, C.byref(d)))
        return int(d.value)

    def BatchSearch(self, queries, k: int, ef: int = 0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        if q.shape[1] != self.Dim():     # the C side reads nq * dim floats: never hand it a shorter buffer
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.Dim(), q.shape[1]))
        nq = q.shape[0]
        ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        _lib.check(_lib.lib().coltt_b200_hnsw_search(self._h, q.ctypes.data_as(_f32p), nq, int(k), int(ef), ids.ctypes.data_as(_u64p),
                                                      sc.ctypes.data_as(_f32p), cnt.ctypes.data_as(_i32p)))
        return ids, sc, cnt

    def TrainPQ(self, num_centroids: int = 256, num_sub_vectors: int = 64, trigger_threshold: int = 65536, iterations: int = 0) -> None:
        """Attach a product quantizer (ProductQuantizerParameters, pkg/models/hnsw_common.go:20-32; builder-defined
        arithmetic, parity unpinned): k-means codebooks on the GPU, then every vertex encoded."""
        pr = _lib.PqParams(int(num_centroids), int(num_sub_vectors), int(trigger_threshold))
        _lib.check(_lib.lib().coltt_b200_hnsw_pq_train(self._h, C.byref(pr), int(iterations)))

    def BatchSearchPQ(self, queries, k: int, ef: int = 0, rerank: bool = True):
        """Hnsw.Search with asymmetric distances over the PQ codes (+ exact re-scoring of the ef survivors)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        q = q.reshape(1, -1) if q.ndim == 1 else q
        if q.shape[1] != self.Dim():
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.Dim(), q.shape[1]))
        nq = q.shape[0]
        ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        _lib.check(_lib.lib().coltt_b200_hnsw_pq_search(self._h, q.ctypes.data_as(_f32p), nq, int(k), int(ef), 1 if rerank else 0,
                                                         ids.ctypes.data_as(_u64p), sc.ctypes.data_as(_f32p), cnt.ctypes.data_as(_i32p)))
        return ids, sc, cnt

    def Search(self, query, k: int, ef: int = 0) -> List[SearchResultItem]:
        """Hnsw.Search(ctx, query, k): nearest first, ascending Score (hnsw.go:268-275)."""
        ids, sc, cnt = self.BatchSearch(np.asarray(query, dtype=np.float32).reshape(1, -1), k, ef)
        return [SearchResultItem(int(ids[0, i]), float(sc[0, i])) for i in range(cnt[0])]

    def last_stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _lib.check(_lib.lib().coltt_b200_hnsw_last_stats(self._h, C.byref(a), C.byref(b)))
        ms = C.c_float(0.0)
        _lib.check(_lib.lib().coltt_b200_hnsw_last_timing(self._h, C.byref(ms)))
        return {"dist_evals": int(a.value), "expansions": int(b.value), "kernel_ms": float(ms.value)}
