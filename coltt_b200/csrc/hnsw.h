// hnsw.h — the device-resident core/vectorindex.Hnsw object shared by hnsw.cu (Load / Search) and
// hnsw_build.cu (bulk construction / Commit).  Internal header.
#pragma once
#include <algorithm>
#include <mutex>
#include <vector>

#include "store.h"

namespace coltt {

struct Hnsw {
  int device = 0, metric = 0, n_sms = 148;
  uint32_t dim = 0, row_stride = 0, n = 0, entry = 0;
  int32_t ef_default = 20, ef_construction = 200, m = 16, m_max = 16, m_max0 = 32, search_algo = 0;
  uint32_t level_mult_bits = 0;                 // hnswConfig.levelMultiplier as stored in the Commit header
  size_t n_edges = 0;
  int32_t max_level = 0;
  uint8_t* d_rows = nullptr; float* d_norm2 = nullptr; uint64_t* d_ids = nullptr; int32_t* d_level = nullptr;
  uint32_t *d_vbase = nullptr, *d_edge_off = nullptr, *d_edge_nbr = nullptr;
  uint32_t* d_edge_dist = nullptr;              // fp32 bits of every edge's stored distance (Commit only)
  uint32_t* d_nbr0 = nullptr;                   // level-0 neighbours [n][nbr0_stride], 0xffffffff-terminated (search only)
  uint32_t nbr0_stride = 32;
  unsigned long long* d_stats = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // around the search kernel(s) of the last call
  float last_kernel_ms = 0.0f;
  std::mutex mu;
  DeviceBuf q_in, q_deq, q_n2, visited, out, counts, q_map, qlog;
  // product quantizer (pq.cu): codebooks [m][c][dsub] fp32, codes [n][m] u8, per-search survivor scratch
  DeviceBuf pq_cent, pq_codes, pq_slots, pq_d2, pq_cnt;
  uint32_t pq_m = 0, pq_c = 0, pq_dsub = 0;
  uint64_t last_evals = 0, last_exp = 0, last_ties = 0;   // last search: distance evaluations, expansions, queries that met a tie
  uint64_t build_fast_queries = 0, build_fast_fallbacks = 0;   // bulk build: searches served by the FAST path / re-run exactly
  double build_ms[4] = {0, 0, 0, 0};            // bulk build: ingest, kNN search, exact edge distances, host graph assembly
  ~Hnsw() {
    cudaSetDevice(device);
    for (void* ptr : {(void*)d_rows, (void*)d_norm2, (void*)d_ids, (void*)d_level, (void*)d_vbase, (void*)d_edge_off, (void*)d_edge_nbr,
                      (void*)d_edge_dist, (void*)d_nbr0, (void*)d_stats})
      if (ptr) cudaFree(ptr);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct HnswEdge {
  uint64_t id;         // neighbour id
  uint32_t slot;       // neighbour slot
  uint32_t dist_bits;  // fp32 bits of the edge distance
};

// pkg/sharding/shard.go:34-41 ShardVertex: FNV-1a over the little-endian id bytes, mod c.
inline uint64_t shard_vertex(uint64_t x, uint64_t c) {
  uint64_t hh = 14695981039346656037ull;
  for (int i = 0; i < 8; i++) {
    hh ^= (x >> (8 * i)) & 0xff;
    hh *= 1099511628211ull;
  }
  return hh % c;
}

template <class T>
inline int upload(T** dst, const std::vector<T>& v) {
  const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  COLTT_CUDA(cudaMalloc((void**)dst, bytes));
  if (!v.empty()) COLTT_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return COLTT_OK;
}

int hnsw_install_graph(Hnsw* h, const std::vector<uint32_t>& vbase, std::vector<std::vector<HnswEdge>>& lists);
int launch_norm2_stored_f32(const uint8_t* rows, uint32_t row_stride, uint32_t dim, size_t n, float* norm2, cudaStream_t stream);

}  // namespace coltt
