"""Host-side mirror of the reference's CFLAT multi-vector collection (experimental/multi_vector_vertex.go):
one fp32 vector per named field and vertex, searched by a weighted sum of per-field scores.  Every field is a
VectorSpace on the GPU; this class keeps them in lock step (same upsert / remove sequence => same slot layout), which is
what `coltt_b200_multi_search` requires.  Vertex ids are strings in the reference (`Id string`); the Go shim maps them
to uint64 handles, which is what crosses the C-ABI — here the caller passes the integer."""
import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np

from . import _lib
from .edge import Metadata, VectorSpace, Quantization_None, SELECT_COMPAT, MATH_EXACT

_u64p, _f32p, _i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int32)


@dataclass
class MultiVectorIndex:
    """experimentalproto.MultiVectorIndex: index_name, vector, include_or_not, ratio."""
    IndexName: str
    Vector: Sequence[float]
    IncludeOrNot: bool = True
    Ratio: int = 0


@dataclass
class NearestNeighbor:
    """experimental.NearestNeighbor (multi_priority_queue.go:37-41); Metadata stays with the caller."""
    Id: int
    Score: float


class MultiVectorVertex:
    def __init__(self, collection_name: str, dim: int, distance: int, vector_fields: Sequence[str], device: int = 0,
                 capacity_hint: int = 0, quantization: int = Quantization_None):
        if quantization != Quantization_None:
            raise ValueError("not support quantization type")          # mutli_vecspace.go:63
        self.collectionName = collection_name
        self.dim, self.distance = dim, distance
        self.fields: Dict[str, VectorSpace] = {
            f: VectorSpace(f"{collection_name}/{f}", Metadata(dim, distance, Quantization_None), device, capacity_hint, SELECT_COMPAT, MATH_EXACT)
            for f in vector_fields}

    def close(self):
        for sp in self.fields.values():
            sp.close()

    def Dim(self) -> int:
        return self.dim

    def LoadSize(self) -> int:
        return next(iter(self.fields.values())).LoadSize() if self.fields else 0

    def ChangedVertex(self, Id: int, multi_vectors: Dict[str, Sequence[float]]) -> None:
        """multiVectorVertex.ChangedVertex (multi_vector_vertex.go:60-74): every vector field of the vertex."""
        self.ChangedVertices(np.array([Id], dtype=np.uint64), {k: np.asarray(v, dtype=np.float32).reshape(1, -1) for k, v in multi_vectors.items()})

    def ChangedVertices(self, ids, multi_vectors: Dict[str, np.ndarray]) -> None:
        if set(multi_vectors) != set(self.fields):
            raise ValueError("a vertex must carry every vector field of the collection: %s" % sorted(self.fields))
        for key, vecs in multi_vectors.items():
            vecs = np.asarray(vecs)
            got = vecs.shape[-1]
            if got != self.dim:
                raise ValueError("index [%s] expect dimension: [%d], but got [%d]" % (key, self.dim, got))
        for key in self.fields:                                         # same order of operations on every field store
            self.fields[key].ChangedVertices(ids, multi_vectors[key])

    def RemoveVertex(self, Id: int) -> None:
        for sp in self.fields.values():
            sp.RemoveVertex(np.array([Id], dtype=np.uint64))

    def MultiVertexSearch(self, topK: int, multi_vectors: List[MultiVectorIndex]) -> List[NearestNeighbor]:
        """multiVectorVertex.MultiVertexSearch (multi_vector_vertex.go:85-137) + validateRatio (experimental_analyzer.go:143-154)."""
        inc = []
        for v in multi_vectors:
            if v.IndexName not in self.fields:
                raise ValueError("index [%s] is not defined vector fields" % v.IndexName)
            if len(v.Vector) != self.dim:
                raise ValueError("index [%s] expect dimension: [%d], but got [%d]" % (v.IndexName, self.dim, len(v.Vector)))
            if v.IncludeOrNot:
                inc.append(v)
        if sum(int(v.Ratio) for v in inc) != 100:
            raise ValueError("sum of the ratios must be 100")
        k = int(topK)
        nf = len(inc)
        handles = (C.c_void_p * nf)(*[self.fields[v.IndexName]._h for v in inc])
        qs = [np.ascontiguousarray(v.Vector, dtype=np.float32) for v in inc]
        qptr = (_f32p * nf)(*[q.ctypes.data_as(_f32p) for q in qs])
        ratios = np.array([int(v.Ratio) for v in inc], dtype=np.int32)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        cnt = C.c_int32(0)
        _lib.check(_lib.lib().coltt_b200_multi_search(handles, qptr, ratios.ctypes.data_as(_i32p), nf, k, ids.ctypes.data_as(_u64p),
                                                      sc.ctypes.data_as(_f32p), C.byref(cnt)))
        return [NearestNeighbor(int(ids[i]), float(sc[i])) for i in range(cnt.value)]
