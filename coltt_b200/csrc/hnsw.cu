// hnsw.cu — K4: device-resident core/vectorindex HNSW search (+ the Commit-blob loader).
//
// Replaces, for search:  Hnsw.Load    core/vectorindex/hnsw_commit.go:164-278
//                        Hnsw.Search  core/vectorindex/hnsw.go:243-278
//                        greedyClosestNeighbor hnsw.go:320-343, searchLevel hnsw.go:345-389,
//                        selectNeighbors hnsw.go:391-397, PriorityQueue core/vectorindex/priority_queue.go
// Data layout in HBM: rows [n][row_stride] fp32 as stored by Insert (already normalized for
// cosine, hnsw.go:105-107), ||row||^2 [n] in AVX lane order, ids [n], level [n], and one CSR
// over (vertex, level): vbase[v] indexes edge_off, neighbours of v at level l are
// edge_nbr[edge_off[vbase[v]+l] .. edge_off[vbase[v]+l+1]) as SLOTS, sorted by neighbour id
// ascending — the reference iterates a Go map (random order); ascending id is the
// deterministic order this build defines, and the walk below is order-exact with respect to it.
//
// One CTA per query.  The traversal is latency/gather bound: per expansion the CTA gathers the
// (<= mMax0) unvisited neighbour rows straight from HBM (2 lanes per row, 4 AVX-lane chains each —
// the same exact arithmetic as flat_scan.cu, so every distance is bit-identical to the Go path),
// then thread 0 replays the reference's sequential heap logic (Go container/heap up/down, restated)
// over the batch.  lowerBound is frozen per expansion in the reference (hnsw.go:357), which is
// what makes the batch legal.  Throughput comes from many resident CTAs (queries) per SM.
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "exact_math.cuh"
#include "hnsw.h"
#include "store.h"

namespace coltt {

static constexpr int kHnswThreads = 128;
static constexpr uint32_t kHnswBatch = 64;       // neighbour rows scored per pass (16 per warp)
static constexpr uint32_t kCandCap = 8192;       // candidate min-heap capacity (shared memory)

struct HnswParams {
  const uint8_t* rows; uint32_t row_stride; uint32_t dim; uint32_t q_stride;
  const float* row_norm2; const uint64_t* ids; const int32_t* level;
  const uint32_t* vbase; const uint32_t* edge_off; const uint32_t* edge_nbr;
  uint32_t n; uint32_t entry; int metric;
  const float* queries; const float* q_norm2;   // prepared (normalized) queries [nq][q_stride]
  uint32_t nq; uint32_t k; uint32_t ef;
  uint32_t* visited;                            // [nq][words] bitmap, zeroed by the caller
  uint32_t visited_words;
  Hit* out; int* out_counts; uint32_t out_stride;
  unsigned long long* stats;                    // [0] distance evaluations, [1] expansions, [2] overflow flag
};

// Go container/heap (src/container/heap/heap.go: up / down), keyed on priority only —
// core/vectorindex/priority_queue.go:160-199: min queue Less = a<b, max queue Less = a>b.
struct SmemHeap {
  float* prio; uint32_t* slot; uint32_t n; bool is_max;
  __device__ __forceinline__ bool less(uint32_t i, uint32_t j) const { return is_max ? prio[i] > prio[j] : prio[i] < prio[j]; }
  __device__ __forceinline__ void swap(uint32_t i, uint32_t j) {
    float p = prio[i]; prio[i] = prio[j]; prio[j] = p;
    uint32_t s = slot[i]; slot[i] = slot[j]; slot[j] = s;
  }
  __device__ void up(uint32_t j) {
    while (j > 0) {
      uint32_t i = (j - 1) / 2;
      if (i == j || !less(j, i)) break;
      swap(i, j);
      j = i;
    }
  }
  __device__ void down(uint32_t i0, uint32_t m) {
    uint32_t i = i0;
    for (;;) {
      uint32_t j1 = 2 * i + 1;
      if (j1 >= m) break;
      uint32_t j = j1, j2 = j1 + 1;
      if (j2 < m && less(j2, j1)) j = j2;
      if (!less(j, i)) break;
      swap(i, j);
      i = j;
    }
  }
  __device__ void push(float p, uint32_t s) { prio[n] = p; slot[n] = s; n++; up(n - 1); }
  __device__ void pop(float& p, uint32_t& s) {
    uint32_t m = n - 1;
    swap(0, m);
    down(0, m);
    p = prio[m]; s = slot[m];
    n = m;
  }
};

template <int METRIC>
__global__ void __launch_bounds__(kHnswThreads) hnsw_search_kernel(HnswParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* q_s = reinterpret_cast<float*>(smem);                       // [q_stride]
  float* cand_p = q_s + p.q_stride;                                  // [kCandCap]
  uint32_t* cand_s = reinterpret_cast<uint32_t*>(cand_p + kCandCap);
  float* res_p = reinterpret_cast<float*>(cand_s + kCandCap);        // [ef+1]
  uint32_t* res_s = reinterpret_cast<uint32_t*>(res_p + (p.ef + 1));
  uint32_t* nb_slot = res_s + (p.ef + 1);                            // [kHnswBatch]
  float* nb_dist = reinterpret_cast<float*>(nb_slot + kHnswBatch);   // [kHnswBatch]
  uint32_t* nb_raw = reinterpret_cast<uint32_t*>(nb_dist + kHnswBatch);  // [kHnswBatch]
  __shared__ uint32_t sh_cur, sh_cnt, sh_state;
  __shared__ float sh_lb, sh_min;
  __shared__ unsigned long long sh_evals, sh_exp;

  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t d = tid; d < p.q_stride; d += blockDim.x) q_s[d] = p.queries[(size_t)q * p.q_stride + d];
  if (tid == 0) { sh_evals = 0; sh_exp = 0; }
  __syncthreads();
  const float qn = METRIC == COLTT_COSINE ? p.q_norm2[q] : 0.0f;
  const uint32_t full8 = (p.dim / 8) * 8;
  const uint32_t r = lane_row16(lane), g = lane_half(lane);
  uint32_t* vis = p.visited + (size_t)q * p.visited_words;

  // exact distances of nb_slot[0..m) -> nb_dist[] (pkg/distance via avx.cpp order; see flat_scan.cu)
  auto batch_dist = [&](uint32_t m) {
    const uint32_t j = warp * 16 + r;
    if (warp * 16 < m) {
      const bool valid = j < m;
      const uint32_t row = valid ? nb_slot[j] : nb_slot[0];
      const uint8_t* rowp = p.rows + (size_t)row * p.row_stride;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (uint32_t e = 0; e < full8; e += 8) {
        float rv[4];
        load4<ELEM_F32>(rowp + (size_t)(e + 4 * g) * 4, nullptr, rv);
        const float4 qv = *reinterpret_cast<const float4*>(q_s + e + 4 * g);
        if (METRIC == COLTT_COSINE) {
          acc[0] = add_rn(acc[0], mul_rn(qv.x, rv[0])); acc[1] = add_rn(acc[1], mul_rn(qv.y, rv[1]));
          acc[2] = add_rn(acc[2], mul_rn(qv.z, rv[2])); acc[3] = add_rn(acc[3], mul_rn(qv.w, rv[3]));
        } else {
          float d0 = sub_rn(qv.x, rv[0]), d1 = sub_rn(qv.y, rv[1]), d2 = sub_rn(qv.z, rv[2]), d3 = sub_rn(qv.w, rv[3]);
          acc[0] = add_rn(acc[0], mul_rn(d0, d0)); acc[1] = add_rn(acc[1], mul_rn(d1, d1));
          acc[2] = add_rn(acc[2], mul_rn(d2, d2)); acc[3] = add_rn(acc[3], mul_rn(d3, d3));
        }
      }
      float h = add_rn(add_rn(acc[0], acc[1]), add_rn(acc[2], acc[3]));
      float o = __shfl_xor_sync(0xffffffffu, h, 8);
      float tot = g == 0 ? add_rn(h, o) : add_rn(o, h);
      for (uint32_t d = full8; d < p.dim; d++) {
        float rv = load1<ELEM_F32>(rowp, d, nullptr), qv = q_s[d];
        if (METRIC == COLTT_COSINE) tot = add_rn(tot, mul_rn(qv, rv));
        else { float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
      }
      if (valid && g == 0) nb_dist[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, p.row_norm2[row]) : sqrt_via_f64(tot);
    }
    __syncthreads();
  };

  // ---- entrypoint distance (hnsw.go:253) and greedy descent through the upper levels (:254-256)
  uint32_t ep = p.entry;
  if (tid == 0) nb_slot[0] = ep;
  __syncthreads();
  batch_dist(1);
  float min_d = nb_dist[0];
  if (tid == 0) sh_evals += 1;
  for (int l = p.level[ep]; l > 0; l--) {
    for (;;) {  // greedyClosestNeighbor, hnsw.go:320-343
      const uint32_t vb = p.vbase[ep];
      const uint32_t e0 = p.edge_off[vb + l], e1 = p.edge_off[vb + l + 1];
      if (tid == 0) { sh_cur = 0xffffffffu; sh_min = min_d; }
      __syncthreads();
      for (uint32_t c0 = e0; c0 < e1; c0 += kHnswBatch) {
        const uint32_t m = e1 - c0 < kHnswBatch ? e1 - c0 : kHnswBatch;
        if (tid < m) nb_slot[tid] = p.edge_nbr[c0 + tid];
        __syncthreads();
        batch_dist(m);
        if (tid == 0) {
          for (uint32_t i = 0; i < m; i++)
            if (nb_dist[i] < sh_min) { sh_min = nb_dist[i]; sh_cur = nb_slot[i]; }
          sh_evals += m;
        }
        __syncthreads();
      }
      const uint32_t closest = sh_cur;
      min_d = sh_min;
      __syncthreads();
      if (closest == 0xffffffffu) break;
      ep = closest;
    }
  }

  // ---- searchLevel(query, ep, ef, 0), hnsw.go:345-389
  SmemHeap cand{cand_p, cand_s, 0, false}, res{res_p, res_s, 0, true};
  if (tid == 0) nb_slot[0] = ep;
  __syncthreads();
  batch_dist(1);                                   // entrypointDistance is recomputed (:346)
  if (tid == 0) {
    sh_evals += 1;
    cand.push(nb_dist[0], ep);
    res.push(nb_dist[0], ep);
    atomicOr(vis + (ep >> 5), 1u << (ep & 31));
    sh_state = 0;
  }
  __syncthreads();
  for (;;) {
    if (tid == 0) {
      if (cand.n == 0) sh_state = 1;
      else {
        float cp; uint32_t cs;
        cand.pop(cp, cs);
        const float lb = res.prio[0];               // resultVertices.Peek() (:357)
        if (cp > lb) sh_state = 1;                  // (:359-361)
        else { sh_cur = cs; sh_lb = lb; sh_exp += 1; }
      }
    }
    __syncthreads();
    if (sh_state) break;
    const uint32_t cur = sh_cur;
    const float lb = sh_lb;
    const uint32_t vb = p.vbase[cur];
    const uint32_t e0 = p.edge_off[vb], e1 = p.edge_off[vb + 1];
    for (uint32_t c0 = e0; c0 < e1; c0 += kHnswBatch) {
      const uint32_t m_raw = e1 - c0 < kHnswBatch ? e1 - c0 : kHnswBatch;
      // visited test-and-set for the whole chunk (:368-371), in parallel; order is restored below
      if (tid < m_raw) {
        const uint32_t s = p.edge_nbr[c0 + tid];
        const uint32_t old = atomicOr(vis + (s >> 5), 1u << (s & 31));
        nb_raw[tid] = (old >> (s & 31)) & 1u ? 0xffffffffu : s;
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < m_raw; i++)
          if (nb_raw[i] != 0xffffffffu) nb_slot[m++] = nb_raw[i];
        sh_cnt = m;
      }
      __syncthreads();
      const uint32_t m = sh_cnt;
      if (m) {
        batch_dist(m);
        if (tid == 0) {
          sh_evals += m;
          for (uint32_t i = 0; i < m; i++) {
            const float d = nb_dist[i];
            if (d < lb || res.n < p.ef) {            // (:374)
              if (cand.n >= kCandCap) { sh_state = 2; break; }
              cand.push(d, nb_slot[i]);
              res.push(d, nb_slot[i]);
              if (res.n > p.ef) { float tp; uint32_t ts; res.pop(tp, ts); }
            }
          }
        }
      }
      __syncthreads();
      if (sh_state == 2) break;
    }
    if (sh_state == 2) break;
  }
  // ---- selectNeighbors(k) (hnsw.go:391-397) and the back-to-front fill (:268-275)
  if (tid == 0) {
    if (sh_state == 2) {
      p.out_counts[q] = 0;
      atomicAdd(p.stats + 2, 1ull);
    } else {
      float tp; uint32_t ts;
      while (res.n > p.k) res.pop(tp, ts);
      const uint32_t n_out = res.n;
      Hit* out = p.out + (size_t)q * p.out_stride;
      for (int i = (int)n_out - 1; i >= 0; i--) {
        res.pop(tp, ts);
        Hit h; h.id = p.ids[ts]; h.score = tp; h.slot = ts;
        out[i] = h;
      }
      p.out_counts[q] = (int)n_out;
    }
    atomicAdd(p.stats + 0, sh_evals);
    atomicAdd(p.stats + 1, sh_exp);
  }
}

// ------------------------------------------------------------------------------------------

struct BlobR {
  const uint8_t* p; size_t n, pos = 0; bool ok = true;
  uint64_t be(int nb) {
    if (pos + nb > n) { ok = false; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v = (v << 8) | p[pos++];
    return v;
  }
  void skip(size_t k) { if (pos + k > n) ok = false; else pos += k; }
};

// CSR over (vertex, level) from per-list edge vectors: neighbours sorted by id (the deterministic iteration
// order), duplicates dropped; edge distances ride along for Commit.
int hnsw_install_graph(Hnsw* h, const std::vector<uint32_t>& vbase, std::vector<std::vector<HnswEdge>>& lists) {
  const uint32_t n = h->n;
  std::vector<uint32_t> edge_off(vbase[n] + 1, 0), edge_nbr;
  std::vector<uint32_t> edge_dist;
  for (uint32_t i = 0; i < vbase[n]; i++) {
    auto& lst = lists[i];
    std::sort(lst.begin(), lst.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id < b.id; });
    lst.erase(std::unique(lst.begin(), lst.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id == b.id; }), lst.end());
    edge_off[i] = (uint32_t)edge_nbr.size();
    for (auto& e : lst) { edge_nbr.push_back(e.slot); edge_dist.push_back(e.dist_bits); }
  }
  edge_off[vbase[n]] = (uint32_t)edge_nbr.size();
  int rc;
  for (void* ptr : {(void*)h->d_vbase, (void*)h->d_edge_off, (void*)h->d_edge_nbr, (void*)h->d_edge_dist})
    if (ptr) cudaFree(ptr);
  h->d_vbase = h->d_edge_off = h->d_edge_nbr = h->d_edge_dist = nullptr;
  if ((rc = upload(&h->d_vbase, vbase)) || (rc = upload(&h->d_edge_off, edge_off)) || (rc = upload(&h->d_edge_nbr, edge_nbr)) ||
      (rc = upload(&h->d_edge_dist, edge_dist)))
    return rc;
  h->n_edges = edge_nbr.size();
  return COLTT_OK;
}

// Hnsw.Load(header=true): hnsw_commit.go:164-278 with hnsw_config.go:203-245 (config) and
// metadata.go:43-105 (per-vertex metadata records are skipped by length).
int hnsw_load(const void* blob, size_t len, int device, Hnsw** out) {
  int rc = require_device(device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(device));
  BlobR r{(const uint8_t*)blob, len};
  std::unique_ptr<Hnsw> h(new Hnsw());
  h->device = device;
  h->search_algo = (int32_t)r.be(4);
  h->level_mult_bits = (uint32_t)r.be(4);  // levelMultiplier (insert-time only; kept for Commit)
  h->ef_default = (int32_t)r.be(4);
  h->ef_construction = (int32_t)r.be(4);
  h->m = (int32_t)r.be(4); h->m_max = (int32_t)r.be(4); h->m_max0 = (int32_t)r.be(4);
  h->dim = (uint32_t)r.be(4);
  const uint8_t di = (uint8_t)r.be(1);
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated HNSW commit header");
  if (di != 1 && di != 2) return fail(COLTT_ERR_FORMAT, "Invalid space type");   // InvalidSpaceTypeErr, hnsw_commit.go:33
  if (h->dim == 0) return fail(COLTT_ERR_FORMAT, "zero dimension in HNSW commit header");
  h->metric = di == 1 ? COLTT_COSINE : COLTT_EUCLIDEAN;
  h->row_stride = (h->dim * 4 + 15) / 16 * 16;
  cudaDeviceProp pr;
  COLTT_CUDA(cudaGetDeviceProperties(&pr, device));
  h->n_sms = pr.multiProcessorCount;
  COLTT_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  COLTT_CUDA(cudaMalloc((void**)&h->d_stats, 3 * sizeof(unsigned long long)));
  std::vector<uint64_t> ids;
  std::vector<int32_t> levels;
  std::vector<uint8_t> rows;
  std::unordered_map<uint64_t, uint32_t> id2slot;
  std::vector<std::vector<uint32_t>> shard_slots(16);
  uint64_t ep_id = 0;
  if (r.pos < len) {  // non-empty index
    ep_id = r.be(8);
    for (int sh = 0; sh < 16 && r.ok; sh++) {
      const uint32_t cnt = (uint32_t)r.be(4);
      for (uint32_t i = 0; i < cnt && r.ok; i++) {
        const uint64_t id = r.be(8);
        const int32_t lvl = (int32_t)r.be(4);
        if (!r.ok || lvl < 0 || lvl > 64) { r.ok = false; break; }
        if (r.pos + (size_t)h->dim * 4 > r.n) { r.ok = false; break; }
        const size_t off = rows.size();
        rows.resize(off + h->row_stride, 0);
        const uint8_t* src = r.p + r.pos;
        for (uint32_t d = 0; d < h->dim; d++)
          for (int b = 0; b < 4; b++) rows[off + (size_t)d * 4 + b] = src[(size_t)d * 4 + (3 - b)];
        r.skip((size_t)h->dim * 4);
        const uint32_t mc = (uint32_t)r.be(2);
        for (uint32_t k = 0; k < mc && r.ok; k++) { r.skip(r.be(1)); r.skip(r.be(2)); }
        const uint32_t slot = (uint32_t)ids.size();
        id2slot[id] = slot;
        ids.push_back(id);
        levels.push_back(lvl);
        shard_slots[sh].push_back(slot);
      }
    }
  }
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated HNSW commit blob (vertices)");
  const uint32_t n = (uint32_t)ids.size();
  std::vector<uint32_t> vbase(n + 1, 0);
  for (uint32_t v = 0; v < n; v++) vbase[v + 1] = vbase[v] + (uint32_t)levels[v] + 1;
  std::vector<std::vector<HnswEdge>> lists(vbase[n]);
  for (int sh = 0; sh < 16 && r.ok; sh++)
    for (size_t i = 0; i < shard_slots[sh].size() && r.ok; i++) {
      const uint64_t id = r.be(8);
      auto it = id2slot.find(id);
      if (!r.ok || it == id2slot.end()) { r.ok = false; break; }
      const uint32_t v = it->second;
      for (int l = levels[v]; l >= 0 && r.ok; l--) {
        const uint32_t ne = (uint32_t)r.be(4);
        auto& lst = lists[vbase[v] + l];
        for (uint32_t j = 0; j < ne && r.ok; j++) {
          const uint64_t nid = r.be(8);
          const uint32_t dbits = (uint32_t)r.be(4);   // stored edge distance: not needed by Search, kept for Commit
          auto nt = id2slot.find(nid);
          if (nt == id2slot.end()) { r.ok = false; break; }
          lst.push_back(HnswEdge{nid, nt->second, dbits});
        }
      }
    }
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated or inconsistent HNSW commit blob (edges)");
  h->n = n;
  for (int32_t l : levels) h->max_level = std::max(h->max_level, l);
  if (n) {
    auto it = id2slot.find(ep_id);
    if (it == id2slot.end()) return fail(COLTT_ERR_FORMAT, "entrypoint id not among the vertices");
    h->entry = it->second;
  }
  if ((rc = upload(&h->d_rows, rows)) || (rc = upload(&h->d_ids, ids)) || (rc = upload(&h->d_level, levels))) return rc;
  if ((rc = hnsw_install_graph(h.get(), vbase, lists))) return rc;
  COLTT_CUDA(cudaMalloc((void**)&h->d_norm2, std::max<size_t>(n, 1) * 4));
  if (n) {
    rc = launch_norm2_stored_f32(h->d_rows, h->row_stride, h->dim, n, h->d_norm2, h->stream);
    if (rc) return rc;
    COLTT_CUDA(cudaStreamSynchronize(h->stream));
  }
  *out = h.release();
  return COLTT_OK;
}

static int hnsw_search(Hnsw* h, const float* queries, size_t nq, int k, int ef_in, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(COLTT_ERR_INVALID, "null argument");
  if (k <= 0) return fail(COLTT_ERR_INVALID, "k must be positive");
  std::lock_guard<std::mutex> lk(h->mu);
  COLTT_CUDA(cudaSetDevice(h->device));
  if (h->n == 0) {  // `if entrypoint == nil { return make(SearchResult, 0), nil }` hnsw.go:249-251
    for (size_t q = 0; q < nq; q++) out_counts[q] = 0;
    return COLTT_OK;
  }
  const uint32_t ef = (uint32_t)std::max(ef_in > 0 ? ef_in : h->ef_default, k);   // gomath.MaxInt(ef, k) hnsw.go:258
  const uint32_t q_stride = (h->dim + 7) / 8 * 8;
  const size_t smem = (size_t)q_stride * 4 + (size_t)kCandCap * 8 + (size_t)(ef + 1) * 8 + kHnswBatch * 12;
  if (smem > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "ef/dim too large for the HNSW kernel's shared memory");
  const uint32_t words = (h->n + 31) / 32;
  cudaStream_t st = h->stream;
  int rc;
  if ((rc = h->q_in.ensure(nq * h->dim * 4)) || (rc = h->q_deq.ensure(nq * q_stride * 4)) || (rc = h->q_n2.ensure(nq * 4)) ||
      (rc = h->visited.ensure(nq * (size_t)words * 4)) || (rc = h->out.ensure(nq * (size_t)k * sizeof(Hit))) || (rc = h->counts.ensure(nq * 4)))
    return rc;
  COLTT_CUDA(cudaMemcpyAsync(h->q_in.p, queries, nq * h->dim * 4, cudaMemcpyHostToDevice, st));
  PrepParams pp{};
  pp.in = (const float*)h->q_in.p; pp.n = nq; pp.in_stride = h->dim; pp.dim = h->dim; pp.smem_stride = (h->dim + 3) / 4 * 4;
  pp.normalize = h->metric == COLTT_COSINE;   // hnsw.go:244-246
  pp.norm2_out = (float*)h->q_n2.p; pp.deq_out = (float*)h->q_deq.p; pp.deq_stride = q_stride;
  rc = launch_prep_rows(pp, ELEM_F32, st);
  if (rc) return rc;
  COLTT_CUDA(cudaMemsetAsync(h->visited.p, 0, nq * (size_t)words * 4, st));
  COLTT_CUDA(cudaMemsetAsync(h->d_stats, 0, 3 * sizeof(unsigned long long), st));
  HnswParams p{};
  p.rows = h->d_rows; p.row_stride = h->row_stride; p.dim = h->dim; p.q_stride = q_stride; p.row_norm2 = h->d_norm2; p.ids = h->d_ids;
  p.level = h->d_level; p.vbase = h->d_vbase; p.edge_off = h->d_edge_off; p.edge_nbr = h->d_edge_nbr; p.n = h->n; p.entry = h->entry;
  p.metric = h->metric; p.queries = (const float*)h->q_deq.p; p.q_norm2 = (const float*)h->q_n2.p; p.nq = (uint32_t)nq; p.k = (uint32_t)k;
  p.ef = ef; p.visited = (uint32_t*)h->visited.p; p.visited_words = words; p.out = (Hit*)h->out.p; p.out_counts = (int*)h->counts.p;
  p.out_stride = (uint32_t)k; p.stats = h->d_stats;
  if (h->metric == COLTT_COSINE) {
    { int arc = kernel_attrs(hnsw_search_kernel<COLTT_COSINE>, smem); if (arc) return arc; }
    hnsw_search_kernel<COLTT_COSINE><<<(unsigned)nq, kHnswThreads, smem, st>>>(p);
  } else {
    { int arc = kernel_attrs(hnsw_search_kernel<COLTT_EUCLIDEAN>, smem); if (arc) return arc; }
    hnsw_search_kernel<COLTT_EUCLIDEAN><<<(unsigned)nq, kHnswThreads, smem, st>>>(p);
  }
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  std::vector<Hit> hits(nq * (size_t)k);
  unsigned long long stats[3];
  COLTT_CUDA(cudaMemcpyAsync(hits.data(), h->out.p, hits.size() * sizeof(Hit), cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaMemcpyAsync(out_counts, h->counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaMemcpyAsync(stats, h->d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  if (stats[2]) return fail(COLTT_ERR_UNSUPPORTED, "HNSW candidate queue overflowed its shared-memory capacity (ef too large)");
  h->last_evals = stats[0];
  h->last_exp = stats[1];
  for (size_t q = 0; q < nq; q++)
    for (int i = 0; i < out_counts[q]; i++) {
      out_ids[q * (size_t)k + i] = hits[q * (size_t)k + i].id;
      out_scores[q * (size_t)k + i] = hits[q * (size_t)k + i].score;
    }
  return COLTT_OK;
}

}  // namespace coltt

using coltt::fail;
using coltt::Hnsw;

extern "C" {
COLTT_API int coltt_b200_hnsw_load(const void* commit_blob, size_t len, int device, coltt_hnsw** out) {
  if (!commit_blob || !out) return fail(COLTT_ERR_INVALID, "null argument");
  Hnsw* h = nullptr;
  int rc = coltt::hnsw_load(commit_blob, len, device, &h);
  if (rc == COLTT_OK) *out = reinterpret_cast<coltt_hnsw*>(h);
  return rc;
}
COLTT_API void coltt_b200_hnsw_destroy(coltt_hnsw* h) { delete reinterpret_cast<Hnsw*>(h); }
COLTT_API int coltt_b200_hnsw_len(coltt_hnsw* h, uint64_t* n) {
  if (!h || !n) return fail(COLTT_ERR_INVALID, "null argument");
  *n = reinterpret_cast<Hnsw*>(h)->n;
  return COLTT_OK;
}
COLTT_API int coltt_b200_hnsw_search(coltt_hnsw* h, const float* queries, size_t nq, int k, int ef, uint64_t* out_ids, float* out_scores,
                                     int32_t* out_counts) {
  if (!h) return fail(COLTT_ERR_INVALID, "null index");
  return coltt::hnsw_search(reinterpret_cast<Hnsw*>(h), queries, nq, k, ef, out_ids, out_scores, out_counts);
}
COLTT_API int coltt_b200_hnsw_last_stats(coltt_hnsw* h, uint64_t* dist_evals, uint64_t* expansions) {
  if (!h || !dist_evals || !expansions) return fail(COLTT_ERR_INVALID, "null argument");
  *dist_evals = reinterpret_cast<Hnsw*>(h)->last_evals;
  *expansions = reinterpret_cast<Hnsw*>(h)->last_exp;
  return COLTT_OK;
}
}
