// store.cu — host side of one edge FLAT collection on one GPU, and its C-ABI.
//
// Mirrors the reference's `vectorspace` implementations (edge/vectorstore.go:30-49):
//   ChangedVertex            -> coltt_b200_store_upsert      (none_vectorstore.go:66-103)
//   RemoveVertex             -> coltt_b200_store_remove      (none_vectorstore.go:105-127)
//   VertexSearch             -> coltt_b200_store_search      (none_vectorstore.go:129-180)
//   FilterableVertexSearch   -> coltt_b200_store_search_subset (none_vectorstore.go:182-253)
//   SaveVertex / LoadVertex  -> coltt_b200_store_export/import (none_vectorstore.go:308-516)
// Data layout in HBM (one allocation each, grown geometrically):
//   rows   [capacity][row_stride]  row-major, element = fp32 / fp16 / f8-compat code,
//                                   row_stride = dim*elem rounded up to 16 B (bulk-copy unit)
//   norm2  [capacity] fp32          ||row||^2 in the AVX lane order (prep.cu)
//   ids    [capacity] u64           slot -> id (ids are sparse snowflakes, edge/id_generator.go)
// Rows are dense: remove moves the last row into the hole.  The id -> slot map lives on the
// host (the Go side owns metadata and filters; it hands ids across the boundary).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "store.h"

namespace coltt {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

int sm100_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; i++) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ok++;
  }
  return ok;
}

int require_device(int device) {
  // cudaGetDeviceProperties costs ~1 ms: validate each ordinal once (hot callers: the per-batch merge)
  static std::atomic<int> ok_cache[64];
  if (device >= 0 && device < 64 && ok_cache[device].load(std::memory_order_relaxed) == 1) return COLTT_OK;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(COLTT_ERR_NO_DEVICE, "no CUDA device visible: libcoltt_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(COLTT_ERR_INVALID, "bad device ordinal");
  cudaDeviceProp pr;
  COLTT_CUDA(cudaGetDeviceProperties(&pr, device));
  if (pr.major != 10) return fail(COLTT_ERR_NO_DEVICE, std::string("device is not sm_100 (Blackwell B200): ") + pr.name);
  if (device < 64) ok_cache[device].store(1, std::memory_order_relaxed);
  return COLTT_OK;
}

static std::atomic<uint64_t> g_alloc_epoch{1};
uint64_t alloc_epoch() { return g_alloc_epoch.load(std::memory_order_relaxed); }

int DeviceBuf::ensure(size_t bytes) {
  if (bytes <= cap) return COLTT_OK;
  g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);   // captured graphs hold the old pointer
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(COLTT_ERR_NOMEM, "cudaMalloc failed for " + std::to_string(bytes) + " bytes");
    }
    want = bytes;
  }
  cap = want;
  return COLTT_OK;
}
DeviceBuf::~DeviceBuf() {
  if (p) cudaFree(p);
}
int PinnedBuf::ensure(size_t bytes) {
  if (bytes <= cap) return COLTT_OK;
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
    return fail(COLTT_ERR_NOMEM, "cudaMallocHost failed");
  }
  cap = bytes;
  return COLTT_OK;
}
// cudaPointerGetAttributes on an ordinary malloc pointer succeeds with cudaMemoryTypeUnregistered (CUDA >= 11); ~1 us.
bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

PinnedBuf::~PinnedBuf() {
  if (p) cudaFreeHost(p);
}

thread_local bool Store::in_fallback = false;
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)(int64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int kernel_attrs(const void* fn, size_t dyn_smem_bytes) {
  struct Key { const void* fn; int dev; bool operator==(const Key& o) const { return fn == o.fn && dev == o.dev; } };
  struct KeyHash { size_t operator()(const Key& k) const { return std::hash<const void*>()(k.fn) ^ ((size_t)k.dev * 0x9e3779b97f4a7c15ull); } };
  static std::mutex mu;
  static std::unordered_map<Key, size_t, KeyHash> seen;   // largest dynamic size configured so far
  int dev = 0;
  COLTT_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  auto it = seen.find(Key{fn, dev});
  if (it == seen.end()) {
    static const bool carve = [] { const char* e = getenv("COLTT_CARVEOUT"); return !e || atoi(e) != 0; }();
    if (carve) COLTT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    it = seen.emplace(Key{fn, dev}, (size_t)48 * 1024).first;   // the limit every kernel has without opting in
  }
  if (dyn_smem_bytes > it->second) {
    COLTT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem_bytes));
    it->second = dyn_smem_bytes;
  }
  return COLTT_OK;
}

SearchCtx::~SearchCtx() {
  for (auto& g : graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
  if (done) cudaEventDestroy(done);
  if (stream) cudaStreamDestroy(stream);
}

__global__ void scatter_ids_kernel(const uint64_t* src, const uint32_t* slots, uint64_t* dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[slots[i]] = src[i];
}

// ------------------------------------------------------------------------------------------
Store::~Store() {
  cudaSetDevice(device);
  pool.clear();
  if (d_rows) cudaFree(d_rows);
  if (d_norm2) cudaFree(d_norm2);
  if (d_scale) cudaFree(d_scale);
  if (d_shadow) cudaFree(d_shadow);
  if (d_ids) cudaFree(d_ids);
  if (d_stat) cudaFree(d_stat);
  if (stream) cudaStreamDestroy(stream);
}

int Store::create(const coltt_store_cfg* cfg, Store** out) {
  if (!cfg || !out) return fail(COLTT_ERR_INVALID, "null argument");
  if (cfg->dim == 0) return fail(COLTT_ERR_INVALID, "dim must be > 0");
  if (cfg->metric != COLTT_COSINE && cfg->metric != COLTT_EUCLIDEAN) return fail(COLTT_ERR_INVALID, "bad metric");
  int elem;
  switch (cfg->quant) {
    case COLTT_QUANT_NONE: elem = ELEM_F32; break;
    case COLTT_QUANT_F16:
    case COLTT_QUANT_BF16: elem = ELEM_F16; break;  // the reference's bf16 IS binary16
    case COLTT_QUANT_F8: elem = ELEM_F8C; break;
    case COLTT_QUANT_F8_E4M3: elem = ELEM_F8E; break;  // builder-defined real fp8: E4M3 codes + one scale per row
    default: return fail(COLTT_ERR_INVALID, "not support quantization type");  // edge/vectorstore.go:78
  }
  int rc = require_device(cfg->device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(cfg->device));
  std::unique_ptr<Store> s(new Store());
  s->cfg = *cfg;
  s->device = cfg->device;
  s->elem = elem;
  s->dim = cfg->dim;
  s->row_stride = (cfg->dim * elem_size(elem) + 15) / 16 * 16;
  // fp32 cosine stores keep an fp16 shadow of their (unit-norm) rows, +50 % memory: COLTT_MATH_FAST filters through it on
  // the tensor cores and re-ranks on the fp32 rows, so results stay bit-identical to EXACT.  L2 rows can leave the fp16
  // range and rows wider than 1536 shadow bytes have no resident-query kernel: those stores are served exactly.
  static const bool shadow_on = [] { const char* e = getenv("COLTT_F32_SHADOW"); return !e || atoi(e) != 0; }();
  if (elem == ELEM_F32 && cfg->metric == COLTT_COSINE && shadow_on && cfg->dim * 2 <= 1536) s->shadow_stride = (cfg->dim * 2 + 15) / 16 * 16;
  cudaDeviceProp pr;
  COLTT_CUDA(cudaGetDeviceProperties(&pr, cfg->device));
  s->n_sms = pr.multiProcessorCount;
  COLTT_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  COLTT_CUDA(cudaMalloc(&s->d_stat, 8));
  COLTT_CUDA(cudaMemset(s->d_stat, 0, 8));
  if (cfg->capacity_hint) {
    rc = s->reserve(cfg->capacity_hint);
    if (rc) return rc;
  }
  *out = s.release();
  return COLTT_OK;
}

int Store::reserve(size_t rows) {
  if (rows <= capacity) return COLTT_OK;
  if (rows > 0xfffffff0ull) return fail(COLTT_ERR_UNSUPPORTED, "more than 2^32 rows per GPU shard");
  size_t nc = std::max<size_t>(rows, std::max<size_t>(capacity * 2, 1024));
  uint8_t *nr = nullptr, *nsh = nullptr;
  float *nn = nullptr, *nsc = nullptr;
  uint64_t* ni = nullptr;
  const bool scaled = elem == ELEM_F8E;
  auto try_alloc = [&](size_t c) {
    nr = nullptr; nn = nullptr; ni = nullptr; nsc = nullptr; nsh = nullptr;
    if (cudaMalloc(&nr, c * row_stride) == cudaSuccess && cudaMalloc(&nn, c * 4) == cudaSuccess &&
        cudaMalloc(&ni, c * 8) == cudaSuccess && (!scaled || cudaMalloc(&nsc, c * 4) == cudaSuccess) &&
        (!shadow_stride || cudaMalloc(&nsh, c * (size_t)shadow_stride) == cudaSuccess))
      return true;
    cudaGetLastError();
    if (nr) cudaFree(nr);
    if (nn) cudaFree(nn);
    if (ni) cudaFree(ni);
    if (nsc) cudaFree(nsc);
    if (nsh) cudaFree(nsh);
    return false;
  };
  if (!try_alloc(nc)) {
    nc = rows;  // geometric growth did not fit: retry with the exact size
    if (!try_alloc(nc)) return fail(COLTT_ERR_NOMEM, "out of device memory growing the store");
  }
  if (n_rows) {
    COLTT_CUDA(cudaMemcpyAsync(nr, d_rows, n_rows * row_stride, cudaMemcpyDeviceToDevice, stream));
    COLTT_CUDA(cudaMemcpyAsync(nn, d_norm2, n_rows * 4, cudaMemcpyDeviceToDevice, stream));
    COLTT_CUDA(cudaMemcpyAsync(ni, d_ids, n_rows * 8, cudaMemcpyDeviceToDevice, stream));
    if (scaled) COLTT_CUDA(cudaMemcpyAsync(nsc, d_scale, n_rows * 4, cudaMemcpyDeviceToDevice, stream));
    if (shadow_stride) COLTT_CUDA(cudaMemcpyAsync(nsh, d_shadow, n_rows * (size_t)shadow_stride, cudaMemcpyDeviceToDevice, stream));
    COLTT_CUDA(cudaStreamSynchronize(stream));
  }
  if (d_rows) cudaFree(d_rows);
  if (d_norm2) cudaFree(d_norm2);
  if (d_scale) cudaFree(d_scale);
  if (d_shadow) cudaFree(d_shadow);
  if (d_ids) cudaFree(d_ids);
  d_rows = nr;
  d_norm2 = nn;
  d_scale = nsc;
  d_shadow = nsh;
  d_ids = ni;
  capacity = nc;
  return COLTT_OK;
}

// ChangedVertex, batched.  Duplicate ids inside one batch: the last one wins, as n sequential
// ChangedVertex calls would leave it.
int Store::upsert(const uint64_t* ids, const float* vecs, size_t n) {
  if (n == 0) return COLTT_OK;
  if (!ids || !vecs) return fail(COLTT_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(mu);
  if (anonymous) return fail(COLTT_ERR_UNSUPPORTED, "store was filled from device memory: search only");
  COLTT_CUDA(cudaSetDevice(device));
  { int wrc = wait_for_searches(); if (wrc) return wrc; }
  std::unordered_map<uint64_t, size_t> last;
  last.reserve(n * 2);
  for (size_t i = 0; i < n; i++) last[ids[i]] = i;
  std::vector<uint32_t> src_idx, slots;
  std::vector<uint64_t> uids;
  src_idx.reserve(last.size()); slots.reserve(last.size()); uids.reserve(last.size());
  size_t new_rows = 0;
  for (size_t i = 0; i < n; i++) {
    if (last[ids[i]] != i) continue;
    auto it = id2slot.find(ids[i]);
    uint32_t slot = it != id2slot.end() ? it->second : (uint32_t)(n_rows + new_rows++);
    src_idx.push_back((uint32_t)i); slots.push_back(slot); uids.push_back(ids[i]);
  }
  int rc = reserve(n_rows + new_rows);
  if (rc) return rc;
  const size_t m = src_idx.size();
  const size_t chunk = std::max<size_t>(1, std::min<size_t>(m, (64u << 20) / ((size_t)dim * 4)));
  rc = up_in.ensure(chunk * dim * 4); if (rc) return rc;
  rc = up_slots.ensure(chunk * 4); if (rc) return rc;
  rc = up_ids.ensure(chunk * 8); if (rc) return rc;
  rc = up_pinned.ensure(chunk * dim * 4); if (rc) return rc;
  for (size_t base = 0; base < m; base += chunk) {
    const size_t c = std::min(chunk, m - base);
    float* stage = (float*)up_pinned.p;
    for (size_t j = 0; j < c; j++) std::memcpy(stage + j * dim, vecs + (size_t)src_idx[base + j] * dim, (size_t)dim * 4);
    COLTT_CUDA(cudaMemcpyAsync(up_in.p, stage, c * dim * 4, cudaMemcpyHostToDevice, stream));
    COLTT_CUDA(cudaMemcpyAsync(up_slots.p, slots.data() + base, c * 4, cudaMemcpyHostToDevice, stream));
    COLTT_CUDA(cudaMemcpyAsync(up_ids.p, uids.data() + base, c * 8, cudaMemcpyHostToDevice, stream));
    PrepParams pp{};
    pp.in = (const float*)up_in.p; pp.n = c; pp.in_stride = dim; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
    pp.normalize = cfg.metric == COLTT_COSINE;  // `if vertex.distance.Type() == T_COSINE` (none_vectorstore.go:96-98)
    pp.rows_out = d_rows; pp.row_stride = row_stride; pp.slots = (const uint32_t*)up_slots.p;
    pp.norm2_out = d_norm2; pp.norm2_by_slot = 1; pp.scale_out = d_scale; pp.shadow_out = d_shadow; pp.shadow_stride = shadow_stride;
    rc = launch_prep_rows(pp, elem, stream);
    if (rc) return rc;
    scatter_ids_kernel<<<(unsigned)((c + 255) / 256), 256, 0, stream>>>((const uint64_t*)up_ids.p, (const uint32_t*)up_slots.p, d_ids, c);
    COLTT_CUDA(cudaGetLastError());
    COLTT_CUDA(cudaStreamSynchronize(stream));  // staging buffers are reused by the next chunk
  }
  h_ids.resize(n_rows + new_rows);
  for (size_t j = 0; j < m; j++) {
    id2slot[uids[j]] = slots[j];
    h_ids[slots[j]] = uids[j];
  }
  n_rows += new_rows;
  return COLTT_OK;
}

__global__ void iota_ids_kernel(uint64_t* ids, size_t base, size_t n, uint64_t id_base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[base + i] = id_base + base + i;
}

// Bulk ingest of rows that already live on the device (the HNSW builder's temporary per-level shards; shards too large
// to stage through host memory): Normalize + Lower + ||row||^2 exactly as upsert does, ids = id_base + slot number.
// The host id map is not maintained, so the store is search-only afterwards.
int Store::append_dev(const float* d_vecs, size_t n, uint32_t stride_floats, uint64_t id_base, bool raw) {
  if (n == 0) return COLTT_OK;
  if (!d_vecs || stride_floats < dim) return fail(COLTT_ERR_INVALID, "append_dev: bad argument");
  std::unique_lock<std::shared_mutex> lk(mu);
  if (n_rows && !anonymous) return fail(COLTT_ERR_UNSUPPORTED, "append_dev on a store with host-mapped ids");
  COLTT_CUDA(cudaSetDevice(device));
  { int wrc = wait_for_searches(); if (wrc) return wrc; }
  int rc = reserve(n_rows + n);
  if (rc) return rc;
  PrepParams pp{};
  pp.in = d_vecs; pp.n = n; pp.in_stride = stride_floats; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
  pp.normalize = cfg.metric == COLTT_COSINE && !raw;
  pp.rows_out = d_rows; pp.row_stride = row_stride; pp.slot_base = (uint32_t)n_rows;
  pp.norm2_out = d_norm2; pp.norm2_by_slot = 1; pp.scale_out = d_scale; pp.shadow_out = d_shadow; pp.shadow_stride = shadow_stride;
  rc = launch_prep_rows(pp, elem, stream);
  if (rc) return rc;
  iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_ids, n_rows, n, id_base);
  COLTT_CUDA(cudaGetLastError());
  COLTT_CUDA(cudaStreamSynchronize(stream));
  n_rows += n;
  anonymous = true;
  return COLTT_OK;
}

int Store::remove(const uint64_t* ids, size_t n) {
  if (n == 0) return COLTT_OK;
  if (!ids) return fail(COLTT_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(mu);
  if (anonymous) return fail(COLTT_ERR_UNSUPPORTED, "store was filled from device memory: search only");
  COLTT_CUDA(cudaSetDevice(device));
  { int wrc = wait_for_searches(); if (wrc) return wrc; }
  for (size_t i = 0; i < n; i++) {
    auto it = id2slot.find(ids[i]);
    if (it == id2slot.end()) continue;  // Go's delete() on a missing key is a no-op
    const uint32_t s = it->second;
    const uint32_t lastslot = (uint32_t)(n_rows - 1);
    id2slot.erase(it);
    if (s != lastslot) {
      COLTT_CUDA(cudaMemcpyAsync(d_rows + (size_t)s * row_stride, d_rows + (size_t)lastslot * row_stride, row_stride, cudaMemcpyDeviceToDevice, stream));
      COLTT_CUDA(cudaMemcpyAsync(d_norm2 + s, d_norm2 + lastslot, 4, cudaMemcpyDeviceToDevice, stream));
      if (d_scale) COLTT_CUDA(cudaMemcpyAsync(d_scale + s, d_scale + lastslot, 4, cudaMemcpyDeviceToDevice, stream));
      if (d_shadow) COLTT_CUDA(cudaMemcpyAsync(d_shadow + (size_t)s * shadow_stride, d_shadow + (size_t)lastslot * shadow_stride, shadow_stride, cudaMemcpyDeviceToDevice, stream));
      COLTT_CUDA(cudaMemcpyAsync(d_ids + s, d_ids + lastslot, 8, cudaMemcpyDeviceToDevice, stream));
      h_ids[s] = h_ids[lastslot];
      id2slot[h_ids[s]] = s;
    }
    h_ids.pop_back();
    n_rows--;
  }
  COLTT_CUDA(cudaStreamSynchronize(stream));
  return COLTT_OK;
}

// Mutations edit rows / norms / ids in place on the store's own stream.  A search enqueued on a caller stream
// (search_dev) may still be running after its call returned and dropped the shared lock, so every mutation, once it
// holds the exclusive lock, first waits for the last work recorded on each pooled search scratch.
int Store::wait_for_searches() {
  std::lock_guard<std::mutex> g(pool_mu);
  for (auto& c : pool)
    if (c->used) COLTT_CUDA(cudaEventSynchronize(c->done));
  return COLTT_OK;
}

// Scratch affinity: work enqueued on one stream may reuse the scratch last used on the same
// stream (stream order protects it); otherwise take a context whose last use has completed.
std::unique_ptr<SearchCtx> Store::acquire_ctx(cudaStream_t user_stream) {
  {
    std::lock_guard<std::mutex> g(pool_mu);
    for (size_t i = 0; i < pool.size(); i++) {
      SearchCtx& c = *pool[i];
      const bool same = user_stream && c.used && c.last_stream == user_stream;
      const bool finished = !c.used || cudaEventQuery(c.done) == cudaSuccess;
      if (same || finished) {
        auto out = std::move(pool[i]);
        pool.erase(pool.begin() + i);
        return out;
      }
    }
    cudaGetLastError();  // cudaErrorNotReady from the queries above is not an error
  }
  std::unique_ptr<SearchCtx> c(new SearchCtx());
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  for (auto& e : c->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return c;
}
void Store::release_ctx(std::unique_ptr<SearchCtx> c) {
  std::lock_guard<std::mutex> g(pool_mu);
  if (c->have_times) { last_ms[0] = c->ms[0]; last_ms[1] = c->ms[1]; last_ms[2] = c->ms[2]; last_ms[3] = c->ms[3]; }
  pool.push_back(std::move(c));
}

template <class Pre, class Post>
int Store::enqueue_cached(SearchCtx& c, cudaStream_t st, const GraphKey& key, bool timed, Pre pre, Post post) {
  static const bool graphs_on = [] { const char* e = getenv("COLTT_GRAPHS"); return !e || atoi(e) != 0; }();
  auto plain = [&]() -> int {
    int rc = pre();
    if (rc) return rc;
    rc = search_enqueue(c, st, (const float*)key.q, key.nq, key.k, key.sel, key.math, nullptr, 0, (Hit*)key.out, (int*)key.counts, timed);
    if (rc) return rc;
    return post();
  };
  if (!graphs_on || timed || key.n_rows == 0) return plain();
  GraphEntry* e = nullptr;
  for (auto& g : c.graphs)
    if (g.key == key) { e = &g; break; }
  if (!e) {
    if (c.graphs.size() >= 16) {
      if (c.graphs.front().exec) cudaGraphExecDestroy(c.graphs.front().exec);
      c.graphs.erase(c.graphs.begin());
    }
    c.graphs.emplace_back();
    e = &c.graphs.back();
    e->key = key;
  }
  if (e->exec && e->epoch != alloc_epoch()) {   // some scratch buffer moved since the capture
    cudaGraphExecDestroy(e->exec);
    e->exec = nullptr;
    e->seen = 1;
  }
  if (e->exec) {
    COLTT_CUDA(cudaGraphLaunch(e->exec, st));
    count_launch((int)e->launches);
    fast_queries += e->fast_q;
    return COLTT_OK;
  }
  if (e->seen <= 0) {            // first sight of this key (or capture is off for it): plain enqueue, which also sizes the scratch
    if (e->seen == 0) e->seen = 1;
    return plain();
  }
  // second call with the same key: capture it
  const uint64_t l0 = launch_count(), f0 = fast_queries, ep0 = alloc_epoch();
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    e->seen = -1;
    return plain();
  }
  int rc = plain();
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &g);
  const uint64_t nl = launch_count() - l0, nf = fast_queries - f0;
  count_launch(-(int)nl);        // nothing ran yet
  fast_queries = f0;
  if (rc || ce != cudaSuccess || !g || ep0 != alloc_epoch()) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    e->seen = -1;
    return plain();              // whatever went wrong under capture is reported (or not) by the plain path
  }
  cudaGraphExec_t ex = nullptr;
  if (cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
    cudaGraphDestroy(g);
    cudaGetLastError();
    e->seen = -1;
    return plain();
  }
  cudaGraphDestroy(g);
  e->exec = ex; e->epoch = ep0; e->launches = nl; e->fast_q = nf;
  COLTT_CUDA(cudaGraphLaunch(ex, st));
  count_launch((int)nl);
  fast_queries += nf;
  return COLTT_OK;
}

// The device part of a search: everything enqueued on `st`; queries and outputs on the device.
// Caller holds the shared lock.
int Store::search_enqueue(SearchCtx& c, cudaStream_t st, const float* d_queries, size_t nq, int k, int select_mode,
                          int math_mode, const uint32_t* d_subset, size_t n_subset, Hit* d_out, int* d_counts,
                          bool timed) {
  if (k <= 0) return fail(COLTT_ERR_INVALID, "top-k must be positive");
  if (select_mode != COLTT_SELECT_COMPAT && select_mode != COLTT_SELECT_NEAREST) return fail(COLTT_ERR_INVALID, "bad select mode");
  if (math_mode != COLTT_MATH_EXACT && math_mode != COLTT_MATH_FAST) return fail(COLTT_ERR_INVALID, "bad math mode");
  const size_t n_items = d_subset ? n_subset : n_rows;
  const int nearest = select_mode == COLTT_SELECT_NEAREST;
  const uint32_t q_stride = (dim + 7) / 8 * 8;
  int rc;
  if (n_items == 0) {
    COLTT_CUDA(cudaMemsetAsync(d_counts, 0, nq * sizeof(int), st));
    if (timed) for (auto& e : c.ev) cudaEventRecord(e, st);
    return COLTT_OK;
  }
  if (math_mode == COLTT_MATH_FAST && !d_subset && !in_fallback) {
    bool used = false;
    rc = fast_enqueue(c, st, d_queries, nq, k, nearest, d_out, d_counts, timed, nullptr, &used);
    if (rc) return rc;
    if (used) return COLTT_OK;   // otherwise the shape is not served by the tensor-core filter: exact path below
  }
  const uint32_t k_eff = (uint32_t)std::min<size_t>((size_t)k, n_items);  // lists never hold more than n_items
  // query batches bounded by list scratch
  ScanPlan plan;
  size_t qb = nq;
  for (;;) {
    rc = plan_flat_scan(elem, dim, row_stride, (uint32_t)n_items, (uint32_t)qb, k_eff, n_sms, &plan);
    if (rc) return rc;
    if (plan.warp_list_bytes <= (512u << 20) || qb <= 8) break;
    qb = std::max<size_t>(8, qb / 2);
  }
  rc = c.q_deq.ensure(nq * q_stride * 4); if (rc) return rc;
  rc = c.q_n2.ensure(nq * 4); if (rc) return rc;
  rc = c.warp_lists.ensure(plan.warp_list_bytes); if (rc) return rc;
  rc = c.cta_lists.ensure(plan.cta_list_bytes); if (rc) return rc;
  rc = c.cta_counts.ensure(plan.cta_count_bytes); if (rc) return rc;

  if (timed) cudaEventRecord(c.ev[0], st);
  // Normalize(target) + Lower(target): *_vectorstore.go:131-139
  PrepParams pp{};
  pp.in = d_queries; pp.n = nq; pp.in_stride = dim; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
  pp.normalize = cfg.metric == COLTT_COSINE && !raw_queries;
  pp.norm2_out = (float*)c.q_n2.p; pp.norm2_by_slot = 0;
  pp.deq_out = (float*)c.q_deq.p; pp.deq_stride = q_stride;
  rc = launch_prep_rows(pp, elem, st);
  if (rc) return rc;
  if (timed) cudaEventRecord(c.ev[1], st);

  for (size_t q0 = 0; q0 < nq; q0 += qb) {
    const size_t nqb = std::min(qb, nq - q0);
    if (nqb != qb) {
      rc = plan_flat_scan(elem, dim, row_stride, (uint32_t)n_items, (uint32_t)nqb, k_eff, n_sms, &plan);
      if (rc) return rc;
    }
    ScanParams sp{};
    sp.rows = d_rows; sp.row_stride = row_stride; sp.dim = dim; sp.n_items = (uint32_t)n_items; sp.subset = d_subset;
    sp.row_norm2 = d_norm2; sp.row_scale = d_scale; sp.ids = d_ids;
    sp.queries = (const float*)c.q_deq.p + q0 * q_stride; sp.q_norm2 = (const float*)c.q_n2.p + q0; sp.q_stride = q_stride;
    sp.nq = (uint32_t)nqb; sp.k = k_eff; sp.nearest = nearest; sp.metric = cfg.metric;
    sp.warp_lists = (Hit*)c.warp_lists.p; sp.cta_lists = (Hit*)c.cta_lists.p; sp.cta_counts = (int*)c.cta_counts.p;
    rc = launch_flat_scan(sp, plan, elem, st);
    if (rc) return rc;
    if (timed && q0 + qb >= nq) cudaEventRecord(c.ev[2], st);
    MergeParams mp{};
    mp.lists = (const Hit*)c.cta_lists.p; mp.counts = (const int*)c.cta_counts.p; mp.n_lists = plan.grid_x;
    mp.nq = (uint32_t)nqb; mp.k_in = k_eff; mp.k = k_eff; mp.nearest = nearest; mp.in_best_first = 1;
    mp.out = d_out + q0 * (size_t)k; mp.out_counts = d_counts + q0;
    if (k_eff != (uint32_t)k) {
      // rows of the public output are k wide; merge into a compact scratch then spread
      rc = c.tmp_out.ensure(nqb * k_eff * sizeof(Hit)); if (rc) return rc;
      mp.out = (Hit*)c.tmp_out.p;
      rc = launch_merge_topk(mp, st); if (rc) return rc;
      COLTT_CUDA(cudaMemcpy2DAsync(d_out + q0 * (size_t)k, (size_t)k * sizeof(Hit), c.tmp_out.p, (size_t)k_eff * sizeof(Hit),
                                   (size_t)k_eff * sizeof(Hit), nqb, cudaMemcpyDeviceToDevice, st));
    } else {
      rc = launch_merge_topk(mp, st); if (rc) return rc;
    }
  }
  if (timed) cudaEventRecord(c.ev[3], st);
  return COLTT_OK;
}

// COLTT_MATH_FAST: tcgen05 filter (gemm_filter.cu) -> exact re-rank + certificate (rerank.cu) ->
// exact re-run of the (rare) queries whose margin could not be certified.
int Store::fast_enqueue(SearchCtx& c, cudaStream_t st, const float* d_queries, size_t nq, int k, int nearest, Hit* d_out,
                        int* d_counts, bool timed, float* dbg_acc, bool* used_fast) {
  *used_fast = false;
  // fp16 rows (kind::f16), E4M3 rows (kind::f8f6f4; cosine: the row scale folds into the per-row coefficient) and fp32
  // cosine rows through their fp16 shadow (kind::f16 on the shadow, exact re-rank on the fp32 rows)
  const bool fp8 = elem == ELEM_F8E, shadow = elem == ELEM_F32 && d_shadow != nullptr;
  if (!(elem == ELEM_F16 || (fp8 && cfg.metric == COLTT_COSINE) || shadow) || n_rows < 4096 || (size_t)k > n_rows) return COLTT_OK;
  const uint8_t* f_rows = shadow ? d_shadow : d_rows;                 // the filter's operand
  const uint32_t f_stride = shadow ? shadow_stride : row_stride, f_es = shadow ? 2u : elem_size(elem);
  GemmPlan gp;
  if (plan_gemm_filter(dim * f_es, fp8, (uint32_t)nq, (uint32_t)k, n_sms, &gp) != COLTT_OK) return COLTT_OK;  // unsupported shape -> exact
  const uint32_t n_cols = gemm_filter_cols(gp, (uint32_t)n_rows);   // columns that see one query
  if (n_cols < gp.groups) return COLTT_OK;                           // too few columns for the bound scheme -> exact
  const uint32_t q_stride = (dim + 7) / 8 * 8;
  int rc;
  rc = c.q_deq.ensure(nq * q_stride * 4); if (rc) return rc;
  rc = c.q_n2.ensure(nq * 4); if (rc) return rc;
  rc = c.q_f16.ensure(nq * (size_t)gp.q_stride); if (rc) return rc;   // lowered queries: fp16 or E4M3 codes, gp.q_stride bytes each
  rc = c.q_scale.ensure(nq * 4); if (rc) return rc;
  rc = c.g_thr.ensure(nq * 4); if (rc) return rc;
  rc = c.g_thr.ensure(nq * 4); if (rc) return rc;
  rc = c.cand.ensure(nq * (size_t)n_cols * gp.cand_out_cap * sizeof(GemmCand)); if (rc) return rc;
  rc = c.cand_cnt.ensure(nq * (size_t)n_cols * 4); if (rc) return rc;
  rc = c.flags.ensure(nq * 4); if (rc) return rc;
  if (!c.fb_cnt.p) {   // [0] flagged-query count, [1] rerank's arrival counter (zero between launches)
    rc = c.fb_cnt.ensure(16); if (rc) return rc;
    COLTT_CUDA(cudaMemsetAsync(c.fb_cnt.p, 0, 16, st));
  }
  rc = c.fb_q.ensure(nq * 4); if (rc) return rc;
  rc = c.pub.ensure(nq * (size_t)n_cols * 4); if (rc) return rc;
  rc = c.cand_buf.ensure((size_t)gp.grid_x * gp.grid_y * 2 * gp.cand_cap * 128 * sizeof(GemmCand)); if (rc) return rc;
  if (timed) cudaEventRecord(c.ev[0], st);
  PrepParams pp{};
  pp.in = d_queries; pp.n = nq; pp.in_stride = dim; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
  pp.normalize = cfg.metric == COLTT_COSINE && !raw_queries;
  pp.norm2_out = (float*)c.q_n2.p; pp.norm2_by_slot = 0;
  pp.deq_out = (float*)c.q_deq.p; pp.deq_stride = q_stride;
  if (fp8) { pp.code_out = (uint8_t*)c.q_f16.p; pp.code_stride = gp.q_stride; pp.scale_out = (float*)c.q_scale.p; }
  else { pp.f16_out = (__half*)c.q_f16.p; pp.f16_stride = gp.q_stride / 2; }
  // scratch initialisation rides in the prep launch: bound = 0, counts = 0, published values = -NaN ("nothing yet":
  // never greater than anything)
  pp.fill_ptr[0] = (uint32_t*)c.g_thr.p;    pp.fill_words[0] = nq;                    pp.fill_value[0] = 0u;
  pp.fill_ptr[1] = (uint32_t*)c.cand_cnt.p; pp.fill_words[1] = nq * (size_t)n_cols;  pp.fill_value[1] = 0u;
  pp.fill_ptr[2] = (uint32_t*)c.pub.p;      pp.fill_words[2] = nq * (size_t)n_cols;  pp.fill_value[2] = 0xffffffffu;
  rc = launch_prep_rows(pp, elem, st); if (rc) return rc;
  if (timed) cudaEventRecord(c.ev[1], st);
  GemmParams g{};
  g.n_rows = (uint32_t)n_rows; g.dim = dim; g.nq = (uint32_t)nq; g.q_lowered = c.q_f16.p; g.q_stride = gp.q_stride;
  g.row_norm2 = d_norm2; g.row_scale = fp8 ? d_scale : nullptr; g.metric = cfg.metric; g.nearest = nearest; g.g_thr = (uint32_t*)c.g_thr.p;
  g.cand_out = (GemmCand*)c.cand.p; g.cand_cnt = (uint32_t*)c.cand_cnt.p; g.dbg_acc = dbg_acc; g.pub = (float*)c.pub.p; g.cand_buf = (GemmCand*)c.cand_buf.p;
#if COLTT_K2_PROF
  {   // role timers of the filter kernel: profiling builds only (make EXTRA=-DCOLTT_K2_PROF=1)
    static const char* prof_env = getenv("COLTT_DEBUG_PROF");
    static const char* flags_env = getenv("COLTT_DEBUG_FLAGS");
    g.dbg_flags = flags_env ? (uint32_t)atoi(flags_env) : 0u;
    if (prof_env) {
      rc = c.prof.ensure((size_t)gp.grid_x * gp.grid_y * 16 * 8); if (rc) return rc;
      COLTT_CUDA(cudaMemsetAsync(c.prof.p, 0, (size_t)gp.grid_x * gp.grid_y * 16 * 8, st));
      g.dbg_prof = (unsigned long long*)c.prof.p;
      g.dbg_prof2 = g.dbg_prof + (size_t)gp.grid_x * gp.grid_y * 8;
    }
  }
#endif
  rc = launch_gemm_filter(g, gp, f_rows, f_stride, st, &c.maps); if (rc) return rc;
  if (timed) cudaEventRecord(c.ev[2], st);
  RerankParams r{};
  r.nq = (uint32_t)nq; r.k = (uint32_t)k; r.dim = dim; r.q_stride = q_stride; r.row_stride = row_stride;
  r.grid_x = n_cols; r.cand_cap = gp.cand_out_cap;
  r.max_rows = gp.pub_kth ? 256u : 64u;
  r.eps_rel = shadow ? fast_eps_rel_f32_shadow(dim) : fast_eps_rel(dim);
  r.metric = cfg.metric; r.nearest = nearest; r.elem = elem;
  r.queries = (const float*)c.q_deq.p; r.q_norm2 = (const float*)c.q_n2.p; r.q_scale = fp8 ? (const float*)c.q_scale.p : nullptr;
  r.rows = d_rows; r.row_norm2 = d_norm2; r.row_scale = d_scale; r.ids = d_ids;
  r.cand_in = (const GemmCand*)c.cand.p; r.cand_cnt = (const uint32_t*)c.cand_cnt.p; r.g_thr = (const uint32_t*)c.g_thr.p;
  r.out = d_out; r.out_stride = (uint32_t)k; r.out_counts = d_counts; r.flags = (uint32_t*)c.flags.p;
  r.n_bad = (uint32_t*)c.fb_cnt.p; r.done_ctr = (uint32_t*)c.fb_cnt.p + 1; r.q_map = (uint32_t*)c.fb_q.p; r.stat_fallbacks = d_stat;
  rc = launch_rerank(r, st); if (rc) return rc;
  if (timed) cudaEventRecord(c.ev[3], st);
#if COLTT_K2_PROF
  if (g.dbg_prof) {
    static int printed = 0;
    COLTT_CUDA(cudaStreamSynchronize(st));
    if (printed++ == 3) {   // a warmed-up step
      const size_t nc = (size_t)gp.grid_x * gp.grid_y;
      std::vector<unsigned long long> h(nc * 16);
      COLTT_CUDA(cudaMemcpy(h.data(), c.prof.p, h.size() * 8, cudaMemcpyDeviceToHost));
      const char* names[16] = {"prod_wait_empty", "prod_total", "mma_wait_tempty", "mma_wait_full", "mma_total", "epi_wait_tfull", "epi_total", "epi_bars",
                               "epi_tmem_ld", "epi_hot_math", "epi_bar2+sweep", "n_slow_entries", "final_cnt", "epi_slow_path", "-", "-"};
      for (int k2 = 0; k2 < 14; k2++) {
        double sum = 0, mx = 0;
        for (size_t i = 0; i < nc; i++) {
          const double v = (double)(k2 < 8 ? h[i * 8 + k2] : h[nc * 8 + i * 8 + (k2 - 8)]);
          sum += v; if (v > mx) mx = v;
        }
        fprintf(stderr, "[coltt prof] %-16s avg %.0f max %.0f\n", names[k2], sum / nc, mx);
      }
    }
  }
#endif
  *used_fast = true;
  fast_queries += nq;
  // Certificate check, on the device: the queries rerank.cu could not certify were compacted into a list by
  // its last CTA and are re-run by the exact kernel, which takes its query count from device memory — no host
  // round trip; with nothing flagged the launches below are one empty wave each.  The flagged list is served in
  // windows so that the exact kernel's per-warp lists stay bounded (512 MB) whatever nq and k are.
  {
    const uint32_t k_eff = (uint32_t)k;   // fast path requires k <= n_rows
    ScanPlan plan;
    rc = plan_flat_scan(elem, dim, row_stride, (uint32_t)n_rows, (uint32_t)nq, k_eff, n_sms, &plan); if (rc) return rc;
    const size_t per_q = (size_t)plan.grid_x * 8 * k_eff * sizeof(Hit);
    size_t win = std::max<size_t>(8, (512u << 20) / per_q / 8 * 8);
    if (win > nq) win = nq;
    rc = plan_flat_scan(elem, dim, row_stride, (uint32_t)n_rows, (uint32_t)win, k_eff, n_sms, &plan); if (rc) return rc;
    plan.grid_y = 1;                      // each CTA loops over the (few) flagged query groups
    rc = c.warp_lists.ensure(plan.warp_list_bytes); if (rc) return rc;
    rc = c.cta_lists.ensure(plan.cta_list_bytes); if (rc) return rc;
    rc = c.cta_counts.ensure(plan.cta_count_bytes); if (rc) return rc;
    for (size_t w0 = 0; w0 < nq; w0 += win) {
      ScanParams sp{};
      sp.rows = d_rows; sp.row_stride = row_stride; sp.dim = dim; sp.n_items = (uint32_t)n_rows; sp.subset = nullptr;
      sp.row_norm2 = d_norm2; sp.row_scale = d_scale; sp.ids = d_ids;
      sp.queries = (const float*)c.q_deq.p; sp.q_norm2 = (const float*)c.q_n2.p; sp.q_stride = q_stride;
      sp.nq = (uint32_t)win; sp.k = k_eff; sp.nearest = nearest; sp.metric = cfg.metric;
      sp.warp_lists = (Hit*)c.warp_lists.p; sp.cta_lists = (Hit*)c.cta_lists.p; sp.cta_counts = (int*)c.cta_counts.p;
      sp.q_map = (const uint32_t*)c.fb_q.p; sp.n_active = (const uint32_t*)c.fb_cnt.p; sp.q_base = (uint32_t)w0;
      rc = launch_flat_scan(sp, plan, elem, st); if (rc) return rc;
      MergeParams mp{};
      mp.lists = (const Hit*)c.cta_lists.p; mp.counts = (const int*)c.cta_counts.p; mp.n_lists = plan.grid_x;
      mp.nq = (uint32_t)win; mp.k_in = k_eff; mp.k = k_eff; mp.nearest = nearest; mp.in_best_first = 1;
      mp.out = d_out; mp.out_counts = d_counts; mp.out_stride = (uint32_t)k;
      mp.q_map = (const uint32_t*)c.fb_q.p; mp.n_active = (const uint32_t*)c.fb_cnt.p; mp.q_base = (uint32_t)w0;
      rc = launch_merge_topk(mp, st); if (rc) return rc;
    }
  }
  return COLTT_OK;
}

int Store::search_host(const float* queries, size_t nq, const uint64_t* cand_ids, size_t n_cand, bool use_subset, int k,
                       int select_mode, int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(COLTT_ERR_INVALID, "null argument");
  if (k <= 0) return fail(COLTT_ERR_INVALID, "top-k must be positive");
  std::shared_lock<std::shared_mutex> lk(mu);
  COLTT_CUDA(cudaSetDevice(device));
  auto ctx = acquire_ctx(nullptr);
  if (!ctx) return fail(COLTT_ERR_CUDA, "could not create a search context");
  struct Rel { Store* s; std::unique_ptr<SearchCtx>* c; ~Rel() { s->release_ctx(std::move(*c)); } } rel{this, &ctx};
  SearchCtx& c = *ctx;
  cudaStream_t st = c.stream;
  const bool timed = timing.load(std::memory_order_relaxed);
  int rc;
  // candidate ids -> slots (the Go map lookup `if node, ok := vertices[shard][uid]; ok`,
  // none_vectorstore.go:203); unknown and repeated ids are dropped
  size_t n_sub = 0;
  if (use_subset) {
    std::vector<uint32_t> slots;
    slots.reserve(n_cand);
    for (size_t i = 0; i < n_cand; i++) {
      auto it = id2slot.find(cand_ids[i]);
      if (it != id2slot.end()) slots.push_back(it->second);
    }
    std::sort(slots.begin(), slots.end());
    slots.erase(std::unique(slots.begin(), slots.end()), slots.end());
    n_sub = slots.size();
    rc = c.subset.ensure(std::max<size_t>(n_sub, 1) * 4); if (rc) return rc;
    if (n_sub) COLTT_CUDA(cudaMemcpyAsync(c.subset.p, slots.data(), n_sub * 4, cudaMemcpyHostToDevice, st));
    COLTT_CUDA(cudaStreamSynchronize(st));  // `slots` is pageable and goes out of scope
  }
  rc = c.q_in.ensure(nq * dim * 4); if (rc) return rc;
  rc = c.out.ensure(nq * (size_t)k * sizeof(Hit)); if (rc) return rc;
  rc = c.counts.ensure(nq * 4); if (rc) return rc;
  rc = c.h_q.ensure(nq * dim * 4); if (rc) return rc;
  rc = c.h_out.ensure(nq * (size_t)k * sizeof(Hit) + nq * 4); if (rc) return rc;
  // A caller buffer that is already page-locked (coltt_b200_host_alloc, cudaHostRegister, a pinned torch tensor) is the DMA
  // source itself; anything else is staged through this scratch's pinned buffer first (one host memcpy of nq*dim*4 bytes).
  const bool direct = host_ptr_is_pinned(queries);
  if (!direct) std::memcpy(c.h_q.p, queries, nq * (size_t)dim * 4);
  const void* h_src = direct ? (const void*)queries : (const void*)c.h_q.p;
  Hit* h_hits = (Hit*)c.h_out.p;
  int* h_counts = (int*)((uint8_t*)c.h_out.p + nq * (size_t)k * sizeof(Hit));
  auto h2d = [&]() -> int {
    COLTT_CUDA(cudaMemcpyAsync(c.q_in.p, h_src, nq * (size_t)dim * 4, cudaMemcpyHostToDevice, st));
    return COLTT_OK;
  };
  auto d2h = [&]() -> int {
    COLTT_CUDA(cudaMemcpyAsync(h_counts, c.counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaMemcpyAsync(h_hits, c.out.p, nq * (size_t)k * sizeof(Hit), cudaMemcpyDeviceToHost, st));
    return COLTT_OK;
  };
  if (use_subset) {
    if ((rc = h2d())) return rc;
    rc = search_enqueue(c, st, (const float*)c.q_in.p, nq, k, select_mode, math_mode, (const uint32_t*)c.subset.p, n_sub, (Hit*)c.out.p,
                        (int*)c.counts.p, timed);
    if (rc) return rc;
    if ((rc = d2h())) return rc;
  } else {
    // H2D + search + D2H as one (cached) graph: every pointer involved is this scratch's own.  With a pinned caller buffer
    // the H2D stays outside the graph (its source changes from call to call) and the graph is search + D2H.
    if (direct) {
      if ((rc = h2d())) return rc;
      auto nop = []() -> int { return COLTT_OK; };
      const GraphKey key{c.q_in.p, nq, k, select_mode, math_mode, c.out.p, c.counts.p, st, n_rows, d_rows, 2};
      rc = enqueue_cached(c, st, key, timed, nop, d2h);
    } else {
      const GraphKey key{c.q_in.p, nq, k, select_mode, math_mode, c.out.p, c.counts.p, st, n_rows, d_rows, 1};
      rc = enqueue_cached(c, st, key, timed, h2d, d2h);
    }
    if (rc) return rc;
  }
  COLTT_CUDA(cudaStreamSynchronize(st));
  c.used = true;
  c.last_stream = st;
  cudaEventRecord(c.done, st);
  const bool had_rows = (use_subset ? n_sub : n_rows) != 0;
  if (had_rows && timed) {
    c.have_times = true;
    cudaEventElapsedTime(&c.ms[0], c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&c.ms[1], c.ev[1], c.ev[2]);
    c.ms[2] = 0.0f;
    cudaEventElapsedTime(&c.ms[3], c.ev[2], c.ev[3]);
  }
  for (size_t q = 0; q < nq; q++) {
    const int n = h_counts[q];
    out_counts[q] = n;
    for (int i = 0; i < n; i++) {
      out_ids[q * (size_t)k + i] = h_hits[q * (size_t)k + i].id;
      out_scores[q * (size_t)k + i] = h_hits[q * (size_t)k + i].score;
    }
  }
  return COLTT_OK;
}

int Store::search_dev(const void* d_queries, size_t nq, int k, int select_mode, int math_mode, void* d_out, void* d_counts,
                      void* stream_) {
  if (nq == 0) return COLTT_OK;
  if (!d_queries || !d_out || !d_counts) return fail(COLTT_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> lk(mu);
  COLTT_CUDA(cudaSetDevice(device));
  auto ctx = acquire_ctx((cudaStream_t)stream_);
  if (!ctx) return fail(COLTT_ERR_CUDA, "could not create a search context");
  struct Rel { Store* s; std::unique_ptr<SearchCtx>* c; ~Rel() { s->release_ctx(std::move(*c)); } } rel{this, &ctx};
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
  const bool timed = timing.load(std::memory_order_relaxed);
  // timings of the previous use of this scratch become readable once that work has finished
  if (timed && ctx->timed_last && ctx->used && n_rows && cudaEventQuery(ctx->ev[3]) == cudaSuccess) {
    cudaEventElapsedTime(&ctx->ms[0], ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->ms[1], ctx->ev[1], ctx->ev[2]);
    ctx->ms[2] = 0.0f;
    cudaEventElapsedTime(&ctx->ms[3], ctx->ev[2], ctx->ev[3]);
    ctx->have_times = true;
  }
  cudaGetLastError();
  int rc;
  if (stream_) {
    const GraphKey key{d_queries, nq, k, select_mode, math_mode, d_out, d_counts, st, n_rows, d_rows, 0};
    auto nop = []() -> int { return COLTT_OK; };
    rc = enqueue_cached(*ctx, st, key, timed, nop, nop);
  } else {
    rc = search_enqueue(*ctx, st, (const float*)d_queries, nq, k, select_mode, math_mode, nullptr, 0, (Hit*)d_out, (int*)d_counts, timed);
  }
  ctx->timed_last = timed;
  ctx->used = true;
  ctx->last_stream = st;
  cudaEventRecord(ctx->done, st);
  if (rc) return rc;
  if (!stream_) {  // own stream: synchronous call
    COLTT_CUDA(cudaStreamSynchronize(st));
    if (n_rows && timed) {
      cudaEventElapsedTime(&ctx->ms[0], ctx->ev[0], ctx->ev[1]);
      cudaEventElapsedTime(&ctx->ms[1], ctx->ev[1], ctx->ev[2]);
      ctx->ms[2] = 0.0f;
      cudaEventElapsedTime(&ctx->ms[3], ctx->ev[2], ctx->ev[3]);
      ctx->have_times = true;
    }
  }
  return COLTT_OK;
}

int Store::get_row(uint64_t id, void* out, size_t out_bytes) {
  std::shared_lock<std::shared_mutex> lk(mu);
  auto it = id2slot.find(id);
  if (it == id2slot.end()) return fail(COLTT_ERR_NOT_FOUND, "id not found");
  const size_t need = (size_t)dim * elem_size(elem);
  if (!out || out_bytes < need) return fail(COLTT_ERR_INVALID, "output buffer too small");
  COLTT_CUDA(cudaSetDevice(device));
  COLTT_CUDA(cudaMemcpy(out, d_rows + (size_t)it->second * row_stride, need, cudaMemcpyDeviceToHost));
  return COLTT_OK;
}

}  // namespace coltt

// ================================== C-ABI ==================================================
using coltt::Store;
using coltt::fail;

extern "C" {

COLTT_API const char* coltt_b200_last_error(void) { return coltt::last_error_cstr(); }
COLTT_API const char* coltt_b200_version(void) { return "coltt_b200 0.1 (sm_100a)"; }
COLTT_API int coltt_b200_device_count(void) { return coltt::sm100_device_count(); }

COLTT_API int coltt_b200_store_create(const coltt_store_cfg* cfg, coltt_store** out) {
  Store* s = nullptr;
  int rc = Store::create(cfg, &s);
  if (rc == COLTT_OK) *out = reinterpret_cast<coltt_store*>(s);
  return rc;
}
COLTT_API void coltt_b200_store_destroy(coltt_store* s) { delete reinterpret_cast<Store*>(s); }
COLTT_API int coltt_b200_store_size(coltt_store* s, uint64_t* n_rows) {
  if (!s || !n_rows) return fail(COLTT_ERR_INVALID, "null argument");
  *n_rows = reinterpret_cast<Store*>(s)->size();
  return COLTT_OK;
}
COLTT_API int coltt_b200_store_dim(coltt_store* s, uint32_t* dim) {
  if (!s || !dim) return fail(COLTT_ERR_INVALID, "null argument");
  *dim = reinterpret_cast<Store*>(s)->dim;
  return COLTT_OK;
}
COLTT_API int coltt_b200_store_upsert(coltt_store* s, const uint64_t* ids, const float* vecs, size_t n) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->upsert(ids, vecs, n);
}
COLTT_API int coltt_b200_store_append_dev(coltt_store* s, const void* d_vecs, size_t n, uint32_t stride_floats, uint64_t id_base) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->append_dev((const float*)d_vecs, n, stride_floats, id_base);
}
COLTT_API float coltt_b200_fast_eps_rel(uint32_t dim) { return coltt::fast_eps_rel(dim); }
COLTT_API int coltt_b200_store_remove(coltt_store* s, const uint64_t* ids, size_t n) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->remove(ids, n);
}
COLTT_API int coltt_b200_store_search(coltt_store* s, const float* queries, size_t nq, int k, int select_mode, int math_mode,
                                      uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->search_host(queries, nq, nullptr, 0, false, k, select_mode, math_mode, out_ids, out_scores, out_counts);
}
COLTT_API int coltt_b200_store_search_subset(coltt_store* s, const float* queries, size_t nq, const uint64_t* cand_ids,
                                             size_t n_cand, int k, int select_mode, uint64_t* out_ids, float* out_scores,
                                             int32_t* out_counts) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  if (n_cand && !cand_ids) return fail(COLTT_ERR_INVALID, "null candidate list");
  return reinterpret_cast<Store*>(s)->search_host(queries, nq, cand_ids, n_cand, true, k, select_mode, COLTT_MATH_EXACT, out_ids, out_scores, out_counts);
}
COLTT_API int coltt_b200_store_search_dev(coltt_store* s, const void* d_queries, size_t nq, int k, int select_mode,
                                          int math_mode, void* d_out, void* d_counts, void* stream) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->search_dev(d_queries, nq, k, select_mode, math_mode, d_out, d_counts, stream);
}
COLTT_API int coltt_b200_merge_topk_dev(int device, const void* d_lists, const void* d_list_counts, int n_lists, size_t nq,
                                        int k_in, int k, int select_mode, void* d_out, void* d_out_counts, void* stream) {
  if (!d_lists || !d_list_counts || !d_out || !d_out_counts || n_lists <= 0 || k <= 0 || k_in <= 0) return fail(COLTT_ERR_INVALID, "bad merge arguments");
  int rc = coltt::require_device(device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(device));
  coltt::MergeParams mp{};
  mp.lists = (const coltt::Hit*)d_lists; mp.counts = (const int*)d_list_counts; mp.n_lists = n_lists; mp.nq = (uint32_t)nq;
  mp.k_in = (uint32_t)k_in; mp.k = (uint32_t)k; mp.nearest = select_mode == COLTT_SELECT_NEAREST; mp.in_best_first = 0;
  mp.out = (coltt::Hit*)d_out; mp.out_counts = (int*)d_out_counts;
  rc = coltt::launch_merge_topk(mp, (cudaStream_t)stream);
  if (rc) return rc;
  if (!stream) COLTT_CUDA(cudaStreamSynchronize(0));
  return COLTT_OK;
}
// Variant for one packed message per rank: rank r's block starts at d_packed + r*rank_stride_bytes and holds
// coltt_hit[nq][k_in] followed (at counts_offset_bytes) by int32 counts[nq].
COLTT_API int coltt_b200_merge_topk_dev2(int device, const void* d_packed, int n_lists, size_t nq, int k_in, int k, int select_mode,
                                         size_t rank_stride_bytes, size_t counts_offset_bytes, void* d_out, void* d_out_counts, void* stream) {
  if (!d_packed || !d_out || !d_out_counts || n_lists <= 0 || k <= 0 || k_in <= 0) return fail(COLTT_ERR_INVALID, "bad merge arguments");
  if (rank_stride_bytes % 16 || counts_offset_bytes % 4) return fail(COLTT_ERR_INVALID, "packed merge: misaligned strides");
  int rc = coltt::require_device(device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(device));
  coltt::MergeParams mp{};
  mp.lists = (const coltt::Hit*)d_packed; mp.counts = (const int*)((const uint8_t*)d_packed + counts_offset_bytes); mp.n_lists = n_lists;
  mp.nq = (uint32_t)nq; mp.k_in = (uint32_t)k_in; mp.k = (uint32_t)k; mp.nearest = select_mode == COLTT_SELECT_NEAREST; mp.in_best_first = 0;
  mp.list_stride_hits = rank_stride_bytes / 16; mp.count_stride = rank_stride_bytes / 4;
  mp.out = (coltt::Hit*)d_out; mp.out_counts = (int*)d_out_counts;
  rc = coltt::launch_merge_topk(mp, (cudaStream_t)stream);
  if (rc) return rc;
  if (!stream) COLTT_CUDA(cudaStreamSynchronize(0));
  return COLTT_OK;
}
COLTT_API int coltt_b200_store_export(coltt_store* s, void* buf, size_t* len) {
  if (!s || !len) return fail(COLTT_ERR_INVALID, "null argument");
  return reinterpret_cast<Store*>(s)->export_blob(buf, len);
}
COLTT_API int coltt_b200_store_import(coltt_store* s, const void* buf, size_t len) {
  if (!s || (!buf && len)) return fail(COLTT_ERR_INVALID, "null argument");
  return reinterpret_cast<Store*>(s)->import_blob(buf, len);
}
COLTT_API int coltt_b200_store_get_row(coltt_store* s, uint64_t id, void* out, size_t out_bytes) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  return reinterpret_cast<Store*>(s)->get_row(id, out, out_bytes);
}
COLTT_API int coltt_b200_store_last_timing(coltt_store* s, float* ms, int n) {
  if (!s || !ms) return fail(COLTT_ERR_INVALID, "null argument");
  Store* st = reinterpret_cast<Store*>(s);
  {
    // an asynchronous search_dev leaves its events in the pooled scratch: harvest finished ones
    std::lock_guard<std::mutex> g(st->pool_mu);
    for (auto& c : st->pool) {
      if (c->used && c->timed_last && cudaEventQuery(c->ev[3]) == cudaSuccess && cudaEventQuery(c->ev[0]) == cudaSuccess) {
        float a, b, d;
        if (cudaEventElapsedTime(&a, c->ev[0], c->ev[1]) == cudaSuccess && cudaEventElapsedTime(&b, c->ev[1], c->ev[2]) == cudaSuccess &&
            cudaEventElapsedTime(&d, c->ev[2], c->ev[3]) == cudaSuccess) {
          st->last_ms[0] = a; st->last_ms[1] = b; st->last_ms[2] = 0.0f; st->last_ms[3] = d;
        }
      }
    }
    cudaGetLastError();
  }
  for (int i = 0; i < n && i < 4; i++) ms[i] = st->last_ms[i];
  return COLTT_OK;
}
COLTT_API int coltt_b200_store_set_timing(coltt_store* s, int on) {
  if (!s) return fail(COLTT_ERR_INVALID, "null store");
  reinterpret_cast<Store*>(s)->timing.store(on != 0, std::memory_order_relaxed);
  return COLTT_OK;
}
COLTT_API uint64_t coltt_b200_kernel_launches(void) { return coltt::launch_count(); }

COLTT_API int coltt_b200_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return fail(COLTT_ERR_INVALID, "bad argument");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(COLTT_ERR_NO_DEVICE, "no CUDA device visible: libcoltt_b200 has no CPU fallback");
  }
  if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return fail(COLTT_ERR_NOMEM, "cudaHostAlloc failed");
  }
  return COLTT_OK;
}
COLTT_API void coltt_b200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- test/diagnostic hooks (not part of include/coltt_b200.h) --------------------------------
// Raw tcgen05 accumulators of the FAST filter for host queries: out_acc is a host [nq][n_rows] fp32.
COLTT_API int coltt_b200_debug_fast_scores(coltt_store* s_, const float* queries, size_t nq, int k, int select_mode, float* out_acc,
                                           uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  Store* s = reinterpret_cast<Store*>(s_);
  if (!s || !queries || !out_acc) return fail(COLTT_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> lk(s->mu);
  COLTT_CUDA(cudaSetDevice(s->device));
  auto ctx = s->acquire_ctx(nullptr);
  if (!ctx) return fail(COLTT_ERR_CUDA, "ctx");
  struct Rel { Store* s; std::unique_ptr<coltt::SearchCtx>* c; ~Rel() { s->release_ctx(std::move(*c)); } } rel{s, &ctx};
  coltt::DeviceBuf dq, dacc, dout, dcnt;
  int rc;
  if ((rc = dq.ensure(nq * s->dim * 4)) || (rc = dacc.ensure(nq * s->n_rows * 4)) || (rc = dout.ensure(nq * (size_t)k * 16)) || (rc = dcnt.ensure(nq * 4))) return rc;
  COLTT_CUDA(cudaMemcpy(dq.p, queries, nq * s->dim * 4, cudaMemcpyHostToDevice));
  COLTT_CUDA(cudaMemset(dacc.p, 0xff, nq * s->n_rows * 4));
  bool used = false;
  rc = s->fast_enqueue(*ctx, ctx->stream, (const float*)dq.p, nq, k, select_mode == COLTT_SELECT_NEAREST, (coltt::Hit*)dout.p, (int*)dcnt.p, false,
                       (float*)dacc.p, &used);
  if (rc) return rc;
  if (!used) return fail(COLTT_ERR_UNSUPPORTED, "shape not served by the FAST path");
  COLTT_CUDA(cudaStreamSynchronize(ctx->stream));
  COLTT_CUDA(cudaMemcpy(out_acc, dacc.p, nq * s->n_rows * 4, cudaMemcpyDeviceToHost));
  if (out_ids && out_scores && out_counts) {
    std::vector<coltt::Hit> h(nq * (size_t)k);
    COLTT_CUDA(cudaMemcpy(h.data(), dout.p, h.size() * 16, cudaMemcpyDeviceToHost));
    COLTT_CUDA(cudaMemcpy(out_counts, dcnt.p, nq * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); i++) { out_ids[i] = h[i].id; out_scores[i] = h[i].score; }
  }
  return COLTT_OK;
}
// [0] queries served by the FAST path, [1] of those re-run exactly because the margin was not certified
COLTT_API int coltt_b200_store_fast_stats(coltt_store* s, uint64_t* out2) {
  if (!s || !out2) return fail(COLTT_ERR_INVALID, "null argument");
  Store* st = reinterpret_cast<Store*>(s);
  unsigned long long fb = 0;
  COLTT_CUDA(cudaSetDevice(st->device));
  COLTT_CUDA(cudaMemcpy(&fb, st->d_stat, 8, cudaMemcpyDeviceToHost));   // counted on the device by rerank.cu
  st->fast_fallbacks = fb;
  out2[0] = st->fast_queries;
  out2[1] = st->fast_fallbacks;
  return COLTT_OK;
}

}  // extern "C"
