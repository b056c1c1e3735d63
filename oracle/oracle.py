"""ctypes binding of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under coltt_b200/ does.

The oracle restates the reference's algorithm (see coltt_oracle.cpp, every function cites
the reference file:line it follows).  When oracle/_ref/libcoltt_ref_avx.so exists (the
reference's own pkg/distance/simd/cpp/avx.cpp compiled unmodified) `use_reference_kernels()`
routes the distance arithmetic through the reference's compiled code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "coltt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
    ref_so = os.path.join(_HERE, "_ref", "libcoltt_ref_avx.so")
    if (force or not os.path.exists(ref_so)) and os.path.exists("/root/reference/pkg/distance/simd/cpp/avx.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "ref"])


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_f16bits_to_f32bits.restype = C.c_uint32
        L.orc_f16bits_to_f32bits.argtypes = [C.c_uint16]
        L.orc_f32bits_to_f16bits.restype = C.c_uint16
        L.orc_f32bits_to_f16bits.argtypes = [C.c_uint32]
        L.orc_f8bits_to_f32bits.restype = C.c_uint32
        L.orc_f8bits_to_f32bits.argtypes = [C.c_uint8]
        L.orc_f32bits_to_f8bits.restype = C.c_uint8
        L.orc_f32bits_to_f8bits.argtypes = [C.c_uint32]
        for name, a, b in (("orc_f32_to_f16_array", f32p, u16p), ("orc_f16_to_f32_array", u16p, f32p),
                           ("orc_f32_to_f8_array", f32p, u8p), ("orc_f8_to_f32_array", u8p, f32p)):
            getattr(L, name).argtypes = [a, b, C.c_size_t]
            getattr(L, name).restype = None
        L.orc_e4m3_decode.argtypes = [C.c_uint8]
        L.orc_e4m3_decode.restype = C.c_float
        L.orc_e4m3_encode.argtypes = [C.c_float]
        L.orc_e4m3_encode.restype = C.c_uint8
        L.orc_e4m3_scale.argtypes = [f32p, C.c_size_t]
        L.orc_e4m3_scale.restype = C.c_float
        L.orc_f32_to_e4m3_array.argtypes = [f32p, u8p, C.c_size_t]
        L.orc_f32_to_e4m3_array.restype = C.c_float
        L.orc_e4m3_to_f32_array.argtypes = [u8p, C.c_float, f32p, C.c_size_t]
        L.orc_e4m3_to_f32_array.restype = None
        L.orc_normalize.argtypes = [f32p, C.c_size_t, f32p]
        L.orc_normalize.restype = None
        L.orc_cosine_dot_norm.argtypes = [C.c_size_t, f32p, f32p, f32p, f32p]
        L.orc_cosine_dot_norm.restype = None
        L.orc_l2sq.argtypes = [C.c_size_t, f32p, f32p, f32p]
        L.orc_l2sq.restype = None
        L.orc_norm2_avx_order.argtypes = [C.c_size_t, f32p]
        L.orc_norm2_avx_order.restype = C.c_float
        L.orc_cosine_distance.argtypes = [C.c_size_t, f32p, f32p]
        L.orc_cosine_distance.restype = C.c_float
        L.orc_euclidean_distance.argtypes = [C.c_size_t, f32p, f32p]
        L.orc_euclidean_distance.restype = C.c_float
        L.orc_score_helper.argtypes = [C.c_float, C.c_int]
        L.orc_score_helper.restype = C.c_float
        L.orc_shard_vertex.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_shard_vertex.restype = C.c_uint64
        L.orc_set_ref_kernels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_ref_kernels.restype = None
        L.orc_store_create.argtypes = [C.c_uint32, C.c_int, C.c_int]
        L.orc_store_create.restype = C.c_void_p
        L.orc_store_destroy.argtypes = [C.c_void_p]
        L.orc_store_destroy.restype = None
        L.orc_store_size.argtypes = [C.c_void_p]
        L.orc_store_size.restype = C.c_uint64
        L.orc_store_upsert.argtypes = [C.c_void_p, u64p, f32p, C.c_size_t]
        L.orc_store_upsert.restype = C.c_int
        L.orc_store_remove.argtypes = [C.c_void_p, u64p, C.c_size_t]
        L.orc_store_remove.restype = C.c_int
        L.orc_store_search.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, u64p, f32p]
        L.orc_store_search.restype = C.c_int
        L.orc_store_search_subset.argtypes = [C.c_void_p, f32p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p]
        L.orc_store_search_subset.restype = C.c_int
        L.orc_store_search_total_order.argtypes = [C.c_void_p, f32p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                   C.c_int, u64p, f32p]
        L.orc_store_search_total_order.restype = C.c_int
        L.orc_store_get_row.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_store_get_row.restype = C.c_int
        L.orc_store_save_vertex.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.orc_store_save_vertex.restype = C.c_size_t
        L.orc_resultset_create.argtypes = [C.c_int]
        L.orc_resultset_create.restype = C.c_void_p
        L.orc_resultset_destroy.argtypes = [C.c_void_p]
        L.orc_resultset_add.argtypes = [C.c_void_p, C.c_uint64, C.c_float]
        L.orc_resultset_add.restype = C.c_int
        L.orc_resultset_to_slice.argtypes = [C.c_void_p, u64p, f32p]
        L.orc_resultset_to_slice.restype = C.c_int
        L.orc_compute_recall.argtypes = [u64p, u64p, C.c_int]
        L.orc_compute_recall.restype = C.c_double
        L.orc_hnsw_create.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_hnsw_create.restype = C.c_void_p
        L.orc_hnsw_destroy.argtypes = [C.c_void_p]
        L.orc_hnsw_len.argtypes = [C.c_void_p]
        L.orc_hnsw_len.restype = C.c_uint64
        L.orc_hnsw_set_ef.argtypes = [C.c_void_p, C.c_int]
        L.orc_hnsw_stats.argtypes = [C.c_void_p, u64p, u64p, C.c_int]
        L.orc_hnsw_level_from_uniform.argtypes = [C.c_void_p, C.c_float]
        L.orc_hnsw_level_from_uniform.restype = C.c_int
        L.orc_hnsw_insert.argtypes = [C.c_void_p, C.c_uint64, f32p, C.c_int]
        L.orc_hnsw_insert.restype = C.c_int
        L.orc_hnsw_search.argtypes = [C.c_void_p, f32p, C.c_int, u64p, f32p]
        L.orc_hnsw_search.restype = C.c_int
        L.orc_hnsw_commit.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.orc_hnsw_commit.restype = C.c_size_t
        L.orc_hnsw_load.argtypes = [u8p, C.c_size_t]
        L.orc_hnsw_load.restype = C.c_void_p
        L.orc_hnsw_remove.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_hnsw_remove.restype = C.c_int
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own avx.cpp, compiled unmodified (None if it was never built)."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libcoltt_ref_avx.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_cosine_similarity_dot_norm.argtypes = [C.c_size_t, f32p, f32p, f32p, f32p]
        R.ref_euclidean_distance_squared.argtypes = [C.c_size_t, f32p, f32p, f32p]
        _REF = R
    return _REF


def use_reference_kernels(on: bool = True) -> bool:
    """Route the oracle's dot/norm/L2 arithmetic through oracle/_ref (the reference's code)."""
    R = ref_lib()
    if not on or R is None:
        lib().orc_set_ref_kernels(None, None)
        return False
    lib().orc_set_ref_kernels(C.cast(R.ref_cosine_similarity_dot_norm, C.c_void_p),
                              C.cast(R.ref_euclidean_distance_squared, C.c_void_p))
    return True


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(u64p)


def aligned_f32(n: int, align: int = 32) -> np.ndarray:
    raw = np.zeros(n * 4 + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * 4].view(np.float32)


# ---------------------------------------------------------------- scalar helpers
def normalize(v):
    v, vp = _f32(v)
    out = np.empty_like(v)
    lib().orc_normalize(vp, v.size, out.ctypes.data_as(f32p))
    return out


def f32_to_f16(a):
    a, ap = _f32(a)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().orc_f32_to_f16_array(ap, out.ctypes.data_as(u16p), a.size)
    return out


def f16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    out = np.empty(a.shape, dtype=np.float32)
    lib().orc_f16_to_f32_array(a.ctypes.data_as(u16p), out.ctypes.data_as(f32p), a.size)
    return out


def f32_to_f8(a):
    a, ap = _f32(a)
    out = np.empty(a.shape, dtype=np.uint8)
    lib().orc_f32_to_f8_array(ap, out.ctypes.data_as(u8p), a.size)
    return out


def f8_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty(a.shape, dtype=np.float32)
    lib().orc_f8_to_f32_array(a.ctypes.data_as(u8p), out.ctypes.data_as(f32p), a.size)
    return out


def e4m3_encode(a):
    """Element-wise OCP E4M3 (fn) encode, RNE, saturating — builder-defined F8_E4M3 store (no scaling here)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    L = lib()
    return np.array([L.orc_e4m3_encode(float(x)) for x in a.ravel()], dtype=np.uint8).reshape(a.shape)


def e4m3_decode(c):
    c = np.ascontiguousarray(c, dtype=np.uint8)
    L = lib()
    return np.array([L.orc_e4m3_decode(int(x)) for x in c.ravel()], dtype=np.float32).reshape(c.shape)


def f32_to_e4m3(v):
    """Lower one vector the way the F8_E4M3 store does: (codes, power-of-two scale)."""
    v, vp = _f32(v)
    out = np.empty(v.shape, dtype=np.uint8)
    s = lib().orc_f32_to_e4m3_array(vp, out.ctypes.data_as(u8p), v.size)
    return out, np.float32(s)


def e4m3_to_f32(codes, scale):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.empty(codes.shape, dtype=np.float32)
    lib().orc_e4m3_to_f32_array(codes.ctypes.data_as(u8p), float(scale), out.ctypes.data_as(f32p), codes.size)
    return out


def cosine_distance(a, b):
    a, ap = _f32(a)
    b, bp = _f32(b)
    return float(np.float32(lib().orc_cosine_distance(a.size, ap, bp)))


def euclidean_distance(a, b):
    a, ap = _f32(a)
    b, bp = _f32(b)
    return float(np.float32(lib().orc_euclidean_distance(a.size, ap, bp)))


def norm2_avx_order(a):
    a, ap = _f32(a)
    return np.float32(lib().orc_norm2_avx_order(a.size, ap))


def shard_vertex(x: int, c: int = 16) -> int:
    return int(lib().orc_shard_vertex(x, c))


COSINE, EUCLIDEAN = 0, 1
Q_NONE, Q_F16, Q_F8, Q_BF16 = 0, 1, 2, 3
Q_F8_E4M3 = 16   # builder-defined real-fp8 store (no reference arithmetic: parity unpinned)
COLTT_COMPAT, NEAREST = 0, 1


class FlatStore:
    """Restatement of edge {none,f16,bf16,f8}_vectorstore.go (search path only)."""

    def __init__(self, dim: int, metric: int = COSINE, quant: int = Q_NONE):
        self.dim, self.metric, self.quant = dim, metric, quant
        self._h = lib().orc_store_create(dim, metric, quant)
        if not self._h:
            raise ValueError("bad store config")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_store_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(lib().orc_store_size(self._h))

    def upsert(self, ids, vecs):  # ChangedVertex
        ids, ip = _u64(ids)
        vecs, vp = _f32(vecs)
        assert vecs.size == ids.size * self.dim
        lib().orc_store_upsert(self._h, ip, vp, ids.size)

    def remove(self, ids):
        ids, ip = _u64(ids)
        lib().orc_store_remove(self._h, ip, ids.size)

    def search(self, query, k, high_cpu=False, select_mode=COLTT_COMPAT, n_threads=1):
        """VertexSearch, literal (Go heap) semantics."""
        q, qp = _f32(query)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        n = lib().orc_store_search(self._h, qp, k, int(high_cpu), select_mode, n_threads,
                                   ids.ctypes.data_as(u64p), sc.ctypes.data_as(f32p))
        return ids[:n], sc[:n]

    def search_subset(self, query, cand_ids, k, high_cpu=False, select_mode=COLTT_COMPAT):
        """FilterableVertexSearch given the inverted index's candidate list."""
        q, qp = _f32(query)
        cand, cp = _u64(cand_ids)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        n = lib().orc_store_search_subset(self._h, qp, cp, cand.size, k, int(high_cpu), select_mode,
                                          ids.ctypes.data_as(u64p), sc.ctypes.data_as(f32p))
        return ids[:n], sc[:n]

    def search_total_order(self, query, k, select_mode=COLTT_COMPAT, cand_ids=None, n_threads=8):
        q, qp = _f32(query)
        if cand_ids is None:
            cand, cp, use = np.zeros(0, np.uint64), None, 0
        else:
            cand, cp = _u64(cand_ids)
            use = 1
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        n = lib().orc_store_search_total_order(self._h, qp, cp, cand.size, use, k, select_mode, n_threads,
                                               ids.ctypes.data_as(u64p), sc.ctypes.data_as(f32p))
        return ids[:n], sc[:n]

    def get_row(self, id_: int):
        dt = {Q_NONE: np.float32, Q_F8: np.uint8, Q_F8_E4M3: np.uint8}.get(self.quant, np.uint16)
        out = np.zeros(self.dim, dtype=dt)
        if lib().orc_store_get_row(self._h, id_, out.ctypes.data_as(C.c_void_p)) != 0:
            raise KeyError(id_)
        return out

    def save_vertex(self) -> bytes:
        n = lib().orc_store_save_vertex(self._h, None, 0)
        buf = np.zeros(n, dtype=np.uint8)
        lib().orc_store_save_vertex(self._h, buf.ctypes.data_as(u8p), n)
        return buf.tobytes()


class Hnsw:
    """Restatement of core/vectorindex/hnsw.go (deterministic: ascending-id neighbour order)."""

    def __init__(self, dim=None, metric=COSINE, m=16, ef=20, ef_construction=200, heuristic=False, _handle=None):
        self._h = _handle if _handle else lib().orc_hnsw_create(dim, metric, m, ef, ef_construction, int(heuristic))
        self.dim = dim

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hnsw_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(lib().orc_hnsw_len(self._h))

    def set_ef(self, ef):
        lib().orc_hnsw_set_ef(self._h, ef)

    def level_from_uniform(self, u: float) -> int:
        return int(lib().orc_hnsw_level_from_uniform(self._h, u))

    def insert(self, id_, vec, level):
        v, vp = _f32(vec)
        return lib().orc_hnsw_insert(self._h, id_, vp, level)

    def build(self, ids, vecs, seed=0xC0177):
        """Insert all rows with levels drawn from a seeded uniform stream (hnsw.go:280-282)."""
        rng = np.random.Generator(np.random.Philox(seed))
        us = rng.random(len(ids), dtype=np.float32)
        us = np.maximum(us, np.float32(1e-30))
        vecs = np.ascontiguousarray(vecs, dtype=np.float32)
        for i, id_ in enumerate(ids):
            self.insert(int(id_), vecs[i], self.level_from_uniform(float(us[i])))

    def remove(self, id_):
        return lib().orc_hnsw_remove(self._h, id_)

    def search(self, query, k):
        q, qp = _f32(query)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        n = lib().orc_hnsw_search(self._h, qp, k, ids.ctypes.data_as(u64p), sc.ctypes.data_as(f32p))
        return ids[:n], sc[:n]

    def stats(self, reset=True):
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().orc_hnsw_stats(self._h, C.byref(a), C.byref(b), int(reset))
        return int(a.value), int(b.value)

    def commit(self) -> bytes:
        n = lib().orc_hnsw_commit(self._h, None, 0)
        buf = np.zeros(n, dtype=np.uint8)
        lib().orc_hnsw_commit(self._h, buf.ctypes.data_as(u8p), n)
        return buf.tobytes()

    @staticmethod
    def load(blob: bytes) -> "Hnsw":
        arr = np.frombuffer(blob, dtype=np.uint8)
        h = lib().orc_hnsw_load(arr.ctypes.data_as(u8p), arr.size)
        if not h:
            raise ValueError("bad commit blob")
        return Hnsw(_handle=h)


def multi_search(dim, metric, ids, fields, included, k):
    """experimental MultiVertexSearch (CFLAT): `fields` = {name: [n, dim] fp32 as handed to ChangedVertex}, `included` =
    [(name, query, ratio)] in request order -> (ids, scores), the k largest scores, descending."""
    names = list(fields)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    n = ids.size
    mats = [np.ascontiguousarray(fields[f], dtype=np.float32).reshape(n, dim) for f in names]
    fptr = (f32p * len(names))(*[m.ctypes.data_as(f32p) for m in mats])
    qs = [np.ascontiguousarray(q, dtype=np.float32) for _, q, _ in included]
    qptr = (f32p * len(qs))(*[q.ctypes.data_as(f32p) for q in qs])
    qf = (C.c_int * len(qs))(*[names.index(f) for f, _, _ in included])
    ratios = (C.c_int * len(qs))(*[int(r) for _, _, r in included])
    out_ids = np.zeros(max(k, 1), dtype=np.uint64)
    out_sc = np.zeros(max(k, 1), dtype=np.float32)
    L = lib()
    L.orc_multi_search.restype = C.c_int
    L.orc_multi_search.argtypes = [C.c_uint32, C.c_int, C.c_size_t, u64p, C.c_int, C.POINTER(f32p), C.c_int, C.POINTER(C.c_int),
                                   C.POINTER(f32p), C.POINTER(C.c_int), C.c_int, u64p, f32p]
    cnt = L.orc_multi_search(dim, metric, n, ids.ctypes.data_as(u64p), len(names), fptr, len(qs), qf, qptr, ratios, k,
                             out_ids.ctypes.data_as(u64p), out_sc.ctypes.data_as(f32p))
    return out_ids[:cnt], out_sc[:cnt]


def compute_recall(base_ids, ids, at):
    b, bp = _u64(base_ids)
    i, ip = _u64(ids)
    return float(lib().orc_compute_recall(bp, ip, at))
