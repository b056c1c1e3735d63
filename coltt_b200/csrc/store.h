// store.h — host-side objects behind the C-ABI handles (internal).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace coltt {

const char* last_error_cstr();
int sm100_device_count();
int require_device(int device);
void count_launch(int n = 1);
// Kernel attributes, set once per (kernel, device) instead of on every launch (cudaFuncSetAttribute takes the
// context lock: a dozen calls per search were a measurable part of the host time of a 0.4 ms step):
//  * the dynamic shared-memory limit, raised only when a launch needs more than any before it;
//  * the maximum shared-memory carve-out for every kernel of the pipeline, so that consecutive launches with
//    different dynamic sizes do not make the SMs switch L1/shared configuration between kernels.
int kernel_attrs(const void* fn, size_t dyn_smem_bytes);
template <class F>
inline int kernel_attrs(F* fn, size_t dyn_smem_bytes) { return kernel_attrs(reinterpret_cast<const void*>(fn), dyn_smem_bytes); }
uint64_t launch_count();

struct DeviceBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  ~DeviceBuf();
};
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  ~PinnedBuf();
};
bool host_ptr_is_pinned(const void* p);   // page-locked host memory known to CUDA: usable as a DMA source in place

// A search whose arguments repeat (same buffers, shapes and modes — a serving loop, bench.py) is captured once into a CUDA
// graph and replayed: the five to seven dependent launches of a FAST search then cost one graph launch, which removes
// ~40 us of launch gaps from a 0.45 ms step.  Entries are dropped when any scratch buffer was reallocated since capture.
struct GraphKey {
  const void* q; size_t nq; int k, sel, math; const void* out; const void* counts; cudaStream_t st; size_t n_rows; const void* rows;
  int host;                         // 0 = device in/out, 1 = + H2D from the pinned staging buffer and D2H, 2 = + D2H only
  bool operator==(const GraphKey& o) const {
    return q == o.q && nq == o.nq && k == o.k && sel == o.sel && math == o.math && out == o.out && counts == o.counts && st == o.st &&
           n_rows == o.n_rows && rows == o.rows && host == o.host;
  }
};
struct GraphEntry {
  GraphKey key{};
  uint64_t epoch = 0, launches = 0, fast_q = 0;
  cudaGraphExec_t exec = nullptr;
  int seen = 0;                     // calls with this key so far; -1 = capture failed once, never try again
};
uint64_t alloc_epoch();             // bumped whenever a DeviceBuf reallocates

// Per-search scratch: own stream + buffers, so that searches on one handle run concurrently
// (the reference searches under per-shard RLocks, edge/none_vectorstore.go:137-146).
struct SearchCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done = nullptr;       // recorded after the last enqueue that used this scratch
  cudaStream_t last_stream = nullptr;
  bool used = false, have_times = false, timed_last = false;
  float ms[4] = {0, 0, 0, 0};
  DeviceBuf q_in, q_deq, q_n2, q_f16, q_scale, warp_lists, cta_lists, cta_counts, out, counts, tmp_out, subset, cand, cand_cnt, g_thr, flags, fb_q, fb_out, fb_cnt, pub, prof, cand_buf, multi_acc;
  PinnedBuf h_q, h_out;
  GemmMapCache maps;                // TMA descriptors of the last FAST launch on this scratch
  std::vector<GraphEntry> graphs;   // captured searches of this scratch (at most 16)
  ~SearchCtx();
};

struct Store {
  coltt_store_cfg cfg{};
  int device = 0, elem = 0, n_sms = 148;
  uint32_t dim = 0, row_stride = 0;
  size_t n_rows = 0, capacity = 0;
  uint8_t* d_rows = nullptr;
  float* d_norm2 = nullptr;
  float* d_scale = nullptr;                        // ELEM_F8E only: per-row power-of-two scale
  uint8_t* d_shadow = nullptr;                     // fp32 cosine stores: fp16 copy of the rows, the tensor-core filter's operand
  uint32_t shadow_stride = 0;                      // bytes per shadow row (0 = no shadow)
  uint64_t* d_ids = nullptr;
  unsigned long long* d_stat = nullptr;            // queries the FAST path re-ran exactly (counted on the device)
  std::atomic<bool> timing{false};                 // record per-phase CUDA events around searches (diagnostics)
  std::vector<uint64_t> h_ids;                     // slot -> id
  std::unordered_map<uint64_t, uint32_t> id2slot;  // id -> slot
  std::shared_mutex mu;                            // searches shared, mutations exclusive
  cudaStream_t stream = nullptr;                   // mutation stream
  DeviceBuf up_in, up_slots, up_ids;
  PinnedBuf up_pinned;
  std::mutex pool_mu;
  std::vector<std::unique_ptr<SearchCtx>> pool;
  float last_ms[4] = {0, 0, 0, 0};

  static int create(const coltt_store_cfg* cfg, Store** out);
  ~Store();
  uint64_t size() {
    std::shared_lock<std::shared_mutex> lk(mu);
    return n_rows;
  }
  int reserve(size_t rows);
  int upsert(const uint64_t* ids, const float* vecs, size_t n);
  int remove(const uint64_t* ids, size_t n);
  // rows already on the device, ids = id_base + slot; raw = store the values as given (no Normalize): internal callers only
  int append_dev(const float* d_vecs, size_t n, uint32_t stride_floats, uint64_t id_base = 0, bool raw = false);
  bool raw_queries = false;                        // internal (HNSW builder): queries arrive already normalized
  int wait_for_searches();                         // mutations: order after outstanding asynchronous searches
  bool anonymous = false;                          // filled by append_dev: no host id map, search only
  int search_host(const float* queries, size_t nq, const uint64_t* cand_ids, size_t n_cand, bool use_subset, int k,
                  int select_mode, int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts);
  int search_dev(const void* d_queries, size_t nq, int k, int select_mode, int math_mode, void* d_out, void* d_counts,
                 void* stream);
  // search_enqueue through the scratch's graph cache (falls back to a plain enqueue while a key is new, timing is on, or
  // capture is not possible); `pre` / `post` enqueue the copies around the search for the host-buffer entry point
  template <class Pre, class Post>
  int enqueue_cached(SearchCtx& c, cudaStream_t st, const GraphKey& key, bool timed, Pre pre, Post post);
  int search_enqueue(SearchCtx& c, cudaStream_t st, const float* d_queries, size_t nq, int k, int select_mode,
                     int math_mode, const uint32_t* d_subset, size_t n_subset, Hit* d_out, int* d_counts, bool timed);
  int get_row(uint64_t id, void* out, size_t out_bytes);
  int fast_enqueue(SearchCtx& c, cudaStream_t st, const float* d_queries, size_t nq, int k, int nearest, Hit* d_out,
                   int* d_counts, bool timed, float* dbg_acc, bool* used_fast);
  uint64_t fast_fallbacks = 0, fast_queries = 0;
  static thread_local bool in_fallback;
  int export_blob(void* buf, size_t* len);
  int import_blob(const void* buf, size_t len);
  std::unique_ptr<SearchCtx> acquire_ctx(cudaStream_t user_stream);
  void release_ctx(std::unique_ptr<SearchCtx> c);
};

}  // namespace coltt
