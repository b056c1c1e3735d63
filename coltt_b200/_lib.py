"""ctypes binding of libcoltt_b200.so (the C-ABI declared in include/coltt_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COLTT_B200_LIB") or os.path.join(_HERE, "lib", "libcoltt_b200.so")   # env override: A/B builds
CSRC = os.path.join(_HERE, "csrc")

# every symbol include/coltt_b200.h declares (tests assert the library exports all of them)
ABI_SYMBOLS = [
    "coltt_b200_last_error", "coltt_b200_version", "coltt_b200_device_count",
    "coltt_b200_store_create", "coltt_b200_store_destroy", "coltt_b200_store_size", "coltt_b200_store_dim",
    "coltt_b200_store_upsert", "coltt_b200_store_remove", "coltt_b200_store_search",
    "coltt_b200_store_search_subset", "coltt_b200_store_search_dev", "coltt_b200_merge_topk_dev", "coltt_b200_merge_topk_dev2",
    "coltt_b200_store_export", "coltt_b200_store_import", "coltt_b200_store_get_row",
    "coltt_b200_hnsw_load", "coltt_b200_hnsw_destroy", "coltt_b200_hnsw_len", "coltt_b200_hnsw_search",
    "coltt_b200_hnsw_last_stats", "coltt_b200_hnsw_build", "coltt_b200_hnsw_commit", "coltt_b200_hnsw_build_stats", "coltt_b200_hnsw_build_fast_stats",
    "coltt_b200_store_last_timing", "coltt_b200_store_set_timing", "coltt_b200_kernel_launches", "coltt_b200_host_alloc", "coltt_b200_host_free", "coltt_b200_multi_search", "coltt_b200_hnsw_last_timing",
    "coltt_b200_store_append_dev", "coltt_b200_store_fast_stats", "coltt_b200_fast_eps_rel", "coltt_b200_hnsw_dim",
    "coltt_b200_init", "coltt_b200_shutdown", "coltt_b200_comm_unique_id", "coltt_b200_comm_init_rank", "coltt_b200_comm_destroy",
    "coltt_b200_comm_info", "coltt_b200_comm_exchange_mode", "coltt_b200_sharded_search", "coltt_b200_sharded_search_dev", "coltt_b200_sharded_search_all",
    "coltt_b200_sharded_hnsw_search", "coltt_b200_sharded_hnsw_pq_search", "coltt_b200_hnsw_pq_train", "coltt_b200_hnsw_pq_search",
]


class ColttError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"coltt_b200 error {code}: {msg}")
        self.code = code
        self.message = msg


class StoreCfg(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("metric", C.c_int32), ("quant", C.c_int32), ("device", C.c_int32),
                ("capacity_hint", C.c_uint64)]


class HnswBuildCfg(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("metric", C.c_int32), ("m", C.c_int32), ("ef", C.c_int32), ("ef_construction", C.c_int32),
                ("device", C.c_int32), ("seed", C.c_uint64)]


class PqParams(C.Structure):
    _fields_ = [("num_centroids", C.c_int32), ("num_sub_vectors", C.c_int32), ("trigger_threshold", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("id", C.c_uint64), ("score", C.c_float), ("slot", C.c_uint32)]


_LIB = None


def build_library(force: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-C", CSRC, "-j8"])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libcoltt_b200.so.  Fails loudly when the extension is missing — there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() / make -C coltt_b200/csrc "
                          "(coltt_b200 has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64p, f32p, i32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.coltt_b200_last_error.restype = C.c_char_p
    L.coltt_b200_version.restype = C.c_char_p
    L.coltt_b200_device_count.restype = C.c_int
    L.coltt_b200_store_create.argtypes = [C.POINTER(StoreCfg), C.POINTER(vp)]
    L.coltt_b200_store_destroy.argtypes = [vp]
    L.coltt_b200_store_destroy.restype = None
    L.coltt_b200_store_size.argtypes = [vp, u64p]
    L.coltt_b200_store_dim.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.coltt_b200_store_upsert.argtypes = [vp, u64p, f32p, C.c_size_t]
    L.coltt_b200_store_remove.argtypes = [vp, u64p, C.c_size_t]
    L.coltt_b200_store_search.argtypes = [vp, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_store_search_subset.argtypes = [vp, f32p, C.c_size_t, u64p, C.c_size_t, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_store_search_dev.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.coltt_b200_merge_topk_dev.argtypes = [C.c_int, vp, vp, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.coltt_b200_merge_topk_dev2.argtypes = [C.c_int, vp, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.coltt_b200_store_export.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
    L.coltt_b200_store_import.argtypes = [vp, vp, C.c_size_t]
    L.coltt_b200_store_get_row.argtypes = [vp, C.c_uint64, vp, C.c_size_t]
    L.coltt_b200_hnsw_load.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.coltt_b200_hnsw_destroy.argtypes = [vp]
    L.coltt_b200_hnsw_destroy.restype = None
    L.coltt_b200_hnsw_len.argtypes = [vp, u64p]
    L.coltt_b200_hnsw_search.argtypes = [vp, f32p, C.c_size_t, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_hnsw_last_stats.argtypes = [vp, u64p, u64p]
    L.coltt_b200_hnsw_last_timing.argtypes = [vp, f32p]
    L.coltt_b200_hnsw_build.argtypes = [C.POINTER(HnswBuildCfg), u64p, f32p, i32p, C.c_size_t, C.POINTER(vp)]
    L.coltt_b200_hnsw_commit.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
    L.coltt_b200_hnsw_build_stats.argtypes = [vp, C.POINTER(C.c_double), u64p, i32p]
    L.coltt_b200_hnsw_build_fast_stats.argtypes = [vp, u64p]
    L.coltt_b200_store_last_timing.argtypes = [vp, f32p, C.c_int]
    L.coltt_b200_store_set_timing.argtypes = [vp, C.c_int]
    L.coltt_b200_kernel_launches.restype = C.c_uint64
    L.coltt_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.coltt_b200_host_free.argtypes = [vp]
    L.coltt_b200_host_free.restype = None
    L.coltt_b200_store_append_dev.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint64]
    L.coltt_b200_store_fast_stats.argtypes = [vp, u64p]
    L.coltt_b200_fast_eps_rel.argtypes = [C.c_uint32]
    L.coltt_b200_fast_eps_rel.restype = C.c_float
    L.coltt_b200_hnsw_dim.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.coltt_b200_hnsw_pq_train.argtypes = [vp, C.POINTER(PqParams), C.c_int]
    L.coltt_b200_hnsw_pq_search.argtypes = [vp, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p, i32p]
    ip = C.POINTER(C.c_int)
    L.coltt_b200_init.argtypes = [ip, C.c_int, C.POINTER(vp)]
    L.coltt_b200_shutdown.restype = None
    L.coltt_b200_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.coltt_b200_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.coltt_b200_comm_destroy.argtypes = [vp]
    L.coltt_b200_comm_destroy.restype = None
    L.coltt_b200_comm_info.argtypes = [vp, ip, ip, ip]
    L.coltt_b200_comm_exchange_mode.argtypes = [vp]
    L.coltt_b200_sharded_search.argtypes = [vp, vp, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_sharded_search_dev.argtypes = [vp, vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.coltt_b200_sharded_search_all.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_sharded_hnsw_search.argtypes = [vp, vp, f32p, C.c_size_t, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_sharded_hnsw_pq_search.argtypes = [vp, vp, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, f32p, i32p]
    L.coltt_b200_multi_search.argtypes = [C.POINTER(vp), C.POINTER(f32p), i32p, C.c_int, C.c_int, u64p, f32p, i32p]
    for name in ABI_SYMBOLS:
        f = getattr(L, name)
        if f.restype is C.c_int and name not in ("coltt_b200_device_count",):
            pass
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise ColttError(rc, (lib().coltt_b200_last_error() or b"").decode("utf-8", "replace"))


def pinned_empty(shape, dtype="float32"):
    """A numpy array over page-locked host memory from coltt_b200_host_alloc (freed with the array).  Query batches
    assembled in such a buffer are DMA'd in place by the host-pointer search calls instead of being staged first."""
    import weakref
    import numpy as np
    dt = np.dtype(dtype)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    nbytes = max(1, int(np.prod(shape)) * dt.itemsize)
    L = lib()
    p = C.c_void_p()
    check(L.coltt_b200_host_alloc(nbytes, C.byref(p)))
    raw = (C.c_uint8 * nbytes).from_address(p.value)
    weakref.finalize(raw, L.coltt_b200_host_free, C.c_void_p(p.value))
    arr = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape))).reshape(shape)   # keeps `raw` alive through .base
    return arr
