"""Host-side mirror of the reference's `edge` vector store interfaces, over the C-ABI.

Same names, argument meaning and error behaviour as the Go code it stands in for, so the parity
tests read like the reference's own call sites:

    edge.Vectorstore            edge/vectorstore.go:51-171   -> Vectorstore
    edge.vectorspace            edge/vectorstore.go:30-49    -> VectorSpace (one collection on one GPU)
    edge.SearchResultItem       edge/priority_queue.go:27-31 -> SearchResultItem
    edge.Metadata               edge/metadata.go             -> Metadata (dim / distance / quantization only)
    scoreHelper                 edge/edge_helper.go:143-148  -> score_helper

Metadata maps, the inverted index and filter expressions stay with the Go caller:
`FilterableVertexSearch` here takes the candidate id list `inverted.SearchWithExpression`
(pkg/inverted/search.go:113-119) would have produced.  All arithmetic runs on the GPU through
libcoltt_b200.so; nothing in this module computes a distance.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib

# edgepb.Distance / edgepb.Quantization (idl/proto/v4/edge.proto:69-80)
Distance_Cosine, Distance_Euclidean = 0, 1
Quantization_None, Quantization_F16, Quantization_F8, Quantization_BF16 = 0, 1, 2, 3
Quantization_F8_E4M3 = 16   # builder extension: real fp8 rows (include/coltt_b200.h), not in edge.proto
SELECT_COMPAT, SELECT_NEAREST = 0, 1   # coltt_select
MATH_EXACT, MATH_FAST = 0, 1           # coltt_math

ErrCollectionExists = "collection: %s is already exists"   # edge/constants.go:31
ErrCollectionNotFound = "collection: %s not found"         # edge/constants.go:29

_u64p, _f32p, _i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int32)
_ELEM_DTYPE = {Quantization_None: np.float32, Quantization_F16: np.uint16, Quantization_BF16: np.uint16,
               Quantization_F8: np.uint8, Quantization_F8_E4M3: np.uint8}


@dataclass
class Metadata:
    """The fields of edge.Metadata the vector path reads (Dimensional/Distancer/Quantizationer)."""
    Dim: int
    Distance: int = Distance_Cosine
    Quantization: int = Quantization_None


@dataclass
class SearchResultItem:  # edge/priority_queue.go:27-31
    Id: int
    Score: float
    Metadata: Optional[dict] = field(default=None)


def score_helper(score: np.ndarray, dist: int):
    """scoreHelper (edge/edge_helper.go:143-148): the RPC-facing score mapping, fp32."""
    s = np.asarray(score, dtype=np.float32)
    if dist == Distance_Cosine:
        return ((np.float32(2) - s) / np.float32(2)) * np.float32(100)
    return np.maximum(np.float64(0), (np.float32(100) - s).astype(np.float64)).astype(np.float32)


class VectorSpace:
    """One collection's vectors on one GPU: the `vectorspace` the reference implements four times
    ({none,f16,bf16,f8}_vectorstore.go)."""

    def __init__(self, collection_name: str, metadata: Metadata, device: int = 0, capacity_hint: int = 0,
                 select_mode: int = SELECT_COMPAT, math_mode: int = MATH_EXACT):
        self.collectionName = collection_name
        self.vertexMetadata = metadata
        self.select_mode = select_mode
        self.math_mode = math_mode
        cfg = _lib.StoreCfg(metadata.Dim, metadata.Distance, metadata.Quantization, device, capacity_hint)
        h = C.c_void_p()
        _lib.check(_lib.lib().coltt_b200_store_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().coltt_b200_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the binding module may already be torn down
            pass

    # -- vectorspace getters (edge/vectorstore.go:43-48)
    def Quantization(self) -> int:
        return self.vertexMetadata.Quantization

    def Distance(self) -> int:
        return self.vertexMetadata.Distance

    def Dim(self) -> int:
        return self.vertexMetadata.Dim

    def LoadSize(self) -> int:
        n = C.c_uint64(0)
        _lib.check(_lib.lib().coltt_b200_store_size(self._h, C.byref(n)))
        return int(n.value)

    # -- mutation
    def ChangedVertex(self, updateID: str, Id: int, vector: Sequence[float]) -> None:
        """vectorspace.ChangedVertex (none_vectorstore.go:66-103).  `updateID` (primary-key upsert) is
        resolved to an existing Id by the caller's inverted index; pass the resolved Id."""
        v = np.ascontiguousarray(vector, dtype=np.float32)
        if v.size != self.Dim():
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.Dim(), v.size))
        self.ChangedVertices(np.array([Id], dtype=np.uint64), v.reshape(1, -1))

    def ChangedVertices(self, ids, vectors) -> None:
        """Batched ChangedVertex (bulk ingest; the reference inserts one RPC at a time)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        vecs = np.ascontiguousarray(vectors, dtype=np.float32)
        if vecs.ndim != 2 or vecs.shape[1] != self.Dim():
            got = vecs.shape[1] if vecs.ndim == 2 else vecs.size
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.Dim(), got))
        if vecs.shape[0] != ids.size:
            raise ValueError("ids and vectors disagree on the number of rows")
        _lib.check(_lib.lib().coltt_b200_store_upsert(self._h, ids.ctypes.data_as(_u64p), vecs.ctypes.data_as(_f32p), ids.size))

    def AppendDeviceRows(self, d_ptr: int, n: int, stride_floats: int = 0, id_base: int = 0) -> None:
        """Bulk ChangedVertex of rows already in device memory (coltt_b200_store_append_dev): ids = id_base + slot."""
        _lib.check(_lib.lib().coltt_b200_store_append_dev(self._h, C.c_void_p(d_ptr), n, stride_floats or self.Dim(), id_base))

    def fast_stats(self):
        fs = (C.c_uint64 * 2)()
        _lib.check(_lib.lib().coltt_b200_store_fast_stats(self._h, fs))
        return {"queries": int(fs[0]), "exact_reruns": int(fs[1])}

    def RemoveVertex(self, drop_ids) -> None:
        """vectorspace.RemoveVertex (none_vectorstore.go:105-127) after dropFilter -> ids."""
        ids = np.ascontiguousarray(drop_ids, dtype=np.uint64)
        _lib.check(_lib.lib().coltt_b200_store_remove(self._h, ids.ctypes.data_as(_u64p), ids.size))

    # -- search
    def _search(self, queries, topK, cand, select_mode, math_mode):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        single = q.ndim == 1
        q = q.reshape(1, -1) if single else q
        if q.shape[1] != self.Dim():
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.Dim(), q.shape[1]))
        nq = q.shape[0]
        k = int(topK)
        ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        sm = self.select_mode if select_mode is None else select_mode
        mm = self.math_mode if math_mode is None else math_mode
        L = _lib.lib()
        if cand is None:
            rc = L.coltt_b200_store_search(self._h, q.ctypes.data_as(_f32p), nq, k, sm, mm, ids.ctypes.data_as(_u64p),
                                           sc.ctypes.data_as(_f32p), cnt.ctypes.data_as(_i32p))
        else:
            cand = np.ascontiguousarray(cand, dtype=np.uint64)
            rc = L.coltt_b200_store_search_subset(self._h, q.ctypes.data_as(_f32p), nq, cand.ctypes.data_as(_u64p), cand.size,
                                                  k, sm, ids.ctypes.data_as(_u64p), sc.ctypes.data_as(_f32p),
                                                  cnt.ctypes.data_as(_i32p))
        _lib.check(rc)
        return ids, sc, cnt, single

    def VertexSearch(self, target, topK: int, highCpu: bool = False, select_mode=None, math_mode=None) -> List[SearchResultItem]:
        """vectorspace.VertexSearch (none_vectorstore.go:129-180).  `highCpu` is accepted and ignored
        (it only chooses between 1 and 16 goroutines in the reference)."""
        ids, sc, cnt, _ = self._search(np.asarray(target, dtype=np.float32).reshape(-1), topK, None, select_mode, math_mode)
        return [SearchResultItem(int(ids[0, i]), float(sc[0, i])) for i in range(cnt[0])]

    def FilterableVertexSearch(self, candidate_ids, target, topK: int, highCpu: bool = False, select_mode=None) -> List[SearchResultItem]:
        """vectorspace.FilterableVertexSearch (none_vectorstore.go:182-253) with the inverted index's answer."""
        ids, sc, cnt, _ = self._search(np.asarray(target, dtype=np.float32).reshape(-1), topK, candidate_ids, select_mode, MATH_EXACT)
        return [SearchResultItem(int(ids[0, i]), float(sc[0, i])) for i in range(cnt[0])]

    def BatchVertexSearch(self, targets, topK: int, select_mode=None, math_mode=None, candidate_ids=None):
        """New surface: nq queries in one call -> (ids [nq,k] u64, scores [nq,k] f32, counts [nq])."""
        ids, sc, cnt, _ = self._search(targets, topK, candidate_ids, select_mode, math_mode)
        return ids, sc, cnt

    # -- persistence (SaveVertex / LoadVertex, none_vectorstore.go:308-516)
    def SaveVertex(self) -> bytes:
        n = C.c_size_t(0)
        _lib.check(_lib.lib().coltt_b200_store_export(self._h, None, C.byref(n)))
        buf = (C.c_uint8 * max(n.value, 1))()
        _lib.check(_lib.lib().coltt_b200_store_export(self._h, buf, C.byref(n)))
        return bytes(buf[: n.value])

    def LoadVertex(self, data: bytes) -> None:
        buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
        _lib.check(_lib.lib().coltt_b200_store_import(self._h, buf, len(data)))

    # -- diagnostics
    def stored_row(self, Id: int) -> np.ndarray:
        """ENode.Vector as stored (normalized + lowered), little-endian element bits."""
        out = np.zeros(self.Dim(), dtype=_ELEM_DTYPE[self.Quantization()])
        _lib.check(_lib.lib().coltt_b200_store_get_row(self._h, Id, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_timing(self, on: bool = True) -> None:
        """Record per-phase CUDA events around searches (off by default: they sit between sub-ms kernels)."""
        _lib.check(_lib.lib().coltt_b200_store_set_timing(self._h, 1 if on else 0))

    def last_timing_ms(self):
        """Phase times of the last search made while set_timing(True) was in effect."""
        ms = (C.c_float * 4)()
        _lib.check(_lib.lib().coltt_b200_store_last_timing(self._h, ms, 4))
        return {"prep": ms[0], "scan": ms[1], "rerank": ms[2], "merge": ms[3]}


class Vectorstore:
    """edge.Vectorstore (edge/vectorstore.go:51-171): collection name -> vectorspace."""

    def __init__(self, device: int = 0, select_mode: int = SELECT_COMPAT, math_mode: int = MATH_EXACT):
        self.Space: Dict[str, VectorSpace] = {}
        self.slock = threading.RLock()
        self.device, self.select_mode, self.math_mode = device, select_mode, math_mode

    def CreateCollection(self, collectionName: str, metadata: Metadata) -> None:
        with self.slock:
            if collectionName in self.Space:
                raise KeyError(ErrCollectionExists % collectionName)
            if metadata.Quantization not in (Quantization_None, Quantization_F16, Quantization_F8, Quantization_BF16, Quantization_F8_E4M3):
                raise ValueError("not support quantization type")  # edge/vectorstore.go:78
            self.Space[collectionName] = VectorSpace(collectionName, metadata, self.device, 0, self.select_mode, self.math_mode)

    def DestroySpace(self, collectionName: str) -> None:
        with self.slock:
            sp = self.Space.pop(collectionName, None)
        if sp:
            sp.close()

    def _space(self, name: str) -> VectorSpace:
        try:
            return self.Space[name]
        except KeyError:
            raise KeyError(ErrCollectionNotFound % name)

    def Quantization(self, name):
        return self._space(name).Quantization()

    def Distance(self, name):
        return self._space(name).Distance()

    def Dim(self, name):
        return self._space(name).Dim()

    def LoadSize(self, name):
        return self._space(name).LoadSize()

    def ChangedVertex(self, collectionName, updateID, Id, metadata, vector):
        return self._space(collectionName).ChangedVertex(updateID, Id, vector)

    def RemoveVertex(self, collectionName, drop_ids):
        return self._space(collectionName).RemoveVertex(drop_ids)

    def VertexSearch(self, collectionName, topK, vector, highCpu=False):
        return self._space(collectionName).VertexSearch(vector, int(topK), highCpu)

    def FilterableVertexSearch(self, collectionName, candidate_ids, topK, vector, highCpu=False):
        return self._space(collectionName).FilterableVertexSearch(candidate_ids, vector, int(topK), highCpu)

    def SavedVertex(self, collectionName) -> bytes:
        return self._space(collectionName).SaveVertex()

    def LoadedVertex(self, collectionName, data: bytes) -> None:
        return self._space(collectionName).LoadVertex(data)
