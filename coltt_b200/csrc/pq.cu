// pq.cu — K6: product quantization for the HNSW walk (BASELINE config 5, "HNSW + distancepq (PQ)").
//
// PARITY UNPINNED.  The reference holds no PQ arithmetic: pkg/distancepq contains only float/bit distance helpers and is
// imported by nothing, and the package that did the work (pkg/hnswpq) is absent from the tree.  What survives
// are the parameter names — ProductQuantizerParameters{NumCentroids <= 256, NumSubVectors, TriggerThreshold}
// (pkg/models/hnsw_common.go:20-32) and the values its playground used (32 sub-vectors x 256 centroids at dim 384,
// playground/hnswpq_verification.go:69-73).  This file is therefore a builder-defined PQ in those terms, judged on
// recall against exact search, not on bit parity:
//   * codebooks: the stored (normalized, for cosine) rows are cut into NumSubVectors contiguous sub-vectors; each
//     sub-quantizer is a Lloyd k-means with NumCentroids centroids (squared L2) trained on TriggerThreshold rows
//     sampled at a fixed stride; every row is then encoded as NumSubVectors one-byte centroid indices;
//   * search: the graph walk of Hnsw.Search (greedy descent through the upper levels, beam of ef at level 0,
//     core/vectorindex/hnsw.go:243-389) with asymmetric distance computation in place of the fp32 distance — a per-query
//     table lut[m][c] = ||q_m - centroid[m][c]||^2 in shared memory, a neighbour costs NumSubVectors table lookups over
//     its NumSubVectors code bytes (64 B at dim 768 / 64 sub-vectors instead of a 3 KB row);
//   * the ef survivors are optionally re-scored with the reference's exact fp32 arithmetic (AVX evaluation order, as
//     flat_scan.cu) so that returned scores are true distances and the final order is exact among the survivors.
// Mapping: one CTA (4 warps) per query.  Warp 0 owns the walk: the result set is a sorted array in shared memory whose
// unexpanded members are the candidate queue (valid because no tie-breaking contract exists here); each expansion
// test-and-sets the visited bitmap for the 32 neighbours at once, then all 128 threads score them (4 threads per
// neighbour, 16-byte code loads).  HBM traffic per query ~ expansions x (128 B list + 32 x NumSubVectors B codes).
#include <cstring>
#include <memory>

#include "exact_math.cuh"
#include "hnsw.h"
#include "store.h"

namespace coltt {

static constexpr int kPqThreads = 128;
static constexpr uint32_t kPqMaxDsub = 32;
static constexpr uint32_t kPqNoSlot = 0xffffffffu;
static constexpr uint32_t kPqExpanded = 0x80000000u;

// ---- training ---------------------------------------------------------------------------------------------------------
// sample i of the training set is row (i * n / T): its m-th sub-vector
__device__ __forceinline__ const float* pq_sub(const uint8_t* rows, uint32_t row_stride, size_t row, uint32_t m, uint32_t dsub) {
  return reinterpret_cast<const float*>(rows + row * row_stride) + (size_t)m * dsub;
}

__global__ void pq_init_kernel(const uint8_t* rows, uint32_t row_stride, size_t n, uint32_t T, uint32_t M, uint32_t C, uint32_t dsub, float* cent) {
  // centroid (m, c) starts at the m-th sub-vector of training sample c * T / C (distinct rows)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * C * dsub) return;
  const uint32_t d = (uint32_t)(i % dsub), c = (uint32_t)((i / dsub) % C), m = (uint32_t)(i / ((size_t)dsub * C));
  const size_t sample = (size_t)c * T / C, row = sample * n / T;
  cent[i] = pq_sub(rows, row_stride, row, m, dsub)[d];
}

// nearest centroid of sub-vector m of `row`; consecutive threads = consecutive rows of one m (centroid reads broadcast)
template <bool TRAIN>
__global__ void __launch_bounds__(256) pq_assign_kernel(const uint8_t* rows, uint32_t row_stride, size_t n, uint32_t T, uint32_t M, uint32_t C, uint32_t dsub,
                                                        const float* cent, uint8_t* codes, float* sums, uint32_t* counts) {
  const size_t per_m = TRAIN ? (size_t)T : n;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t m = blockIdx.y;
  if (i >= per_m) return;
  const size_t row = TRAIN ? i * n / T : i;
  const float* x = pq_sub(rows, row_stride, row, m, dsub);
  float xv[kPqMaxDsub];
#pragma unroll
  for (uint32_t d = 0; d < kPqMaxDsub; d++) xv[d] = d < dsub ? x[d] : 0.0f;
  const float* cm = cent + (size_t)m * C * dsub;
  float best = __int_as_float(0x7f800000);
  uint32_t bc = 0;
  for (uint32_t c = 0; c < C; c++) {
    float s = 0.0f;
#pragma unroll
    for (uint32_t d = 0; d < kPqMaxDsub; d++)
      if (d < dsub) { const float df = xv[d] - cm[(size_t)c * dsub + d]; s = fmaf(df, df, s); }
    if (s < best) { best = s; bc = c; }
  }
  if (TRAIN) {
    float* sm = sums + ((size_t)m * C + bc) * dsub;
#pragma unroll
    for (uint32_t d = 0; d < kPqMaxDsub; d++)
      if (d < dsub) atomicAdd(sm + d, xv[d]);
    atomicAdd(counts + (size_t)m * C + bc, 1u);
  } else {
    codes[row * M + m] = (uint8_t)bc;
  }
}

__global__ void pq_update_kernel(uint32_t M, uint32_t C, uint32_t dsub, const float* sums, const uint32_t* counts, float* cent) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * C * dsub) return;
  const uint32_t cnt = counts[i / dsub];
  if (cnt) cent[i] = sums[i] / (float)cnt;      // an empty cluster keeps its centroid
}

// ---- search -----------------------------------------------------------------------------------------------------------
struct PqSearchParams {
  const uint8_t* codes; const float* cent; uint32_t M, C, dsub;
  const float* queries; uint32_t q_stride;       // prepared (normalized for cosine) fp32 queries
  const int32_t* level; const uint32_t *vbase, *edge_off, *edge_nbr, *nbr0; uint32_t nbr0_stride;
  uint32_t n, entry, nq, ef;
  uint32_t* visited; uint32_t visited_words;      // [nq][words], zeroed by the caller
  uint32_t* out_slots; float* out_d2; int* out_counts;   // [nq][ef] survivors, ascending ADC distance
  unsigned long long* stats;                      // [0] code evaluations, [1] expansions
};

__global__ void __launch_bounds__(kPqThreads) pq_search_kernel(PqSearchParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* lut = reinterpret_cast<float*>(smem);                           // [M][C]
  float* res_d = lut + (size_t)p.M * p.C;                                 // [ef]
  uint32_t* res_s = reinterpret_cast<uint32_t*>(res_d + p.ef);            // [ef]  slot | kPqExpanded
  uint32_t* nb_slot = res_s + p.ef;                                       // [32]
  float* nb_d = reinterpret_cast<float*>(nb_slot + 32);                   // [32]
  __shared__ uint32_t sh_m, sh_done, sh_n;
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qv = p.queries + (size_t)q * p.q_stride;

  // asymmetric-distance table: lut[m][c] = ||q_m - centroid[m][c]||^2
  for (uint32_t e = tid; e < p.M * p.C; e += blockDim.x) {
    const uint32_t m = e / p.C;
    const float* c = p.cent + (size_t)e * p.dsub;
    const float* x = qv + (size_t)m * p.dsub;
    float s = 0.0f;
    for (uint32_t d = 0; d < p.dsub; d++) { const float df = x[d] - c[d]; s = fmaf(df, df, s); }
    lut[e] = s;
  }
  if (tid == 0) { sh_n = 0; sh_done = 0; sh_m = 0; }
  __syncthreads();

  uint32_t* vis = p.visited + (size_t)q * p.visited_words;
  unsigned long long evals = 0, exps = 0;
  // score nb_slot[0..m): 4 threads per neighbour, each a quarter of the code bytes
  auto score = [&](uint32_t m) {
    const uint32_t j = tid >> 2, part = tid & 3;
    float s = 0.0f;
    if (j < m) {
      const uint8_t* code = p.codes + (size_t)nb_slot[j] * p.M;
      const uint32_t per = (p.M + 3) / 4, m0 = part * per, m1 = m0 + per < p.M ? m0 + per : p.M;
      if ((per & 15u) == 0 && (p.M & 15u) == 0) {
        for (uint32_t mm = m0; mm < m1; mm += 16) {
          const uint4 w = *reinterpret_cast<const uint4*>(code + mm);
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int b = 0; b < 4; b++) s += lut[(size_t)(mm + 4 * u + b) * p.C + ((ws[u] >> (8 * b)) & 0xffu)];
        }
      } else {
        for (uint32_t mm = m0; mm < m1; mm++) s += lut[(size_t)mm * p.C + code[mm]];
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (j < m && part == 0) nb_d[j] = s;
  };

  // ---- entrypoint + greedy descent through the upper levels (hnsw.go:253-256,320-343) with ADC distances
  uint32_t cur = p.entry;
  if (tid == 0) nb_slot[0] = cur;
  __syncthreads();
  score(1);
  __syncthreads();
  float d_cur = nb_d[0];
  evals += 1;
  for (int l = p.level[cur]; l > 0; l--) {
    for (;;) {
      const uint32_t vb = p.vbase[cur];
      const uint32_t e0 = p.edge_off[vb + l], e1 = p.edge_off[vb + l + 1];
      float best = d_cur;
      uint32_t best_s = kPqNoSlot;
      for (uint32_t c0 = e0; c0 < e1; c0 += 32) {
        const uint32_t m = e1 - c0 < 32 ? e1 - c0 : 32;
        __syncthreads();
        if (tid < m) nb_slot[tid] = p.edge_nbr[c0 + tid];
        __syncthreads();
        score(m);
        __syncthreads();
        evals += m;
        for (uint32_t i = 0; i < m; i++)
          if (nb_d[i] < best) { best = nb_d[i]; best_s = nb_slot[i]; }
      }
      if (best_s == kPqNoSlot) break;
      cur = best_s;
      d_cur = best;
    }
  }

  // ---- level 0: beam of ef (hnsw.go:345-389).  The result set is a sorted array; its unexpanded members are the candidates.
  if (tid == 0) {
    res_d[0] = d_cur; res_s[0] = cur; sh_n = 1;
    atomicOr(vis + (cur >> 5), 1u << (cur & 31));
  }
  __syncthreads();
  for (;;) {
    if (warp == 0) {
      // pick the closest unexpanded member
      const uint32_t n = sh_n;
      uint32_t idx = kPqNoSlot;
      for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t b = __ballot_sync(0xffffffffu, i < n && !(res_s[i] & kPqExpanded));
        if (b) { idx = base + __ffs(b) - 1; break; }
      }
      uint32_t m = 0;
      if (idx == kPqNoSlot) {
        if (lane == 0) sh_done = 1;
      } else {
        const uint32_t c = res_s[idx];
        __syncwarp();
        if (lane == 0) res_s[idx] = c | kPqExpanded;
        exps++;
        // its neighbour list: test-and-set the visited bitmap, keep the fresh ones (list order)
        for (uint32_t off = 0; off < p.nbr0_stride && m < 32; off += 32) {
          const uint32_t s = p.nbr0[(size_t)c * p.nbr0_stride + off + lane];
          bool fresh = false;
          if (s != kPqNoSlot) {
            const uint32_t bit = 1u << (s & 31);
            fresh = !(atomicOr(vis + (s >> 5), bit) & bit);
          }
          const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
          const uint32_t pos = m + __popc(fm & ((1u << lane) - 1u));
          if (fresh && pos < 32) nb_slot[pos] = s;
          // neighbours beyond 32 fresh ones in one expansion (degree > 32 lists only) are un-marked again
          if (fresh && pos >= 32) atomicAnd(vis + (s >> 5), ~(1u << (s & 31)));
          m = min(32u, m + (uint32_t)__popc(fm));
          if (__ballot_sync(0xffffffffu, s == kPqNoSlot)) break;      // lists are kPqNoSlot-terminated
        }
        if (lane == 0) sh_m = m;
      }
    }
    __syncthreads();
    if (sh_done) break;
    const uint32_t m = sh_m;
    score(m);
    evals += m;
    __syncthreads();
    if (warp == 0) {
      // fold the scored neighbours into the sorted result set (hnsw.go:374-382: d < worst || not full)
      for (uint32_t j = 0; j < m; j++) {
        const float d = nb_d[j];
        const uint32_t s = nb_slot[j];
        uint32_t n = sh_n;
        if (n == p.ef && !(d < res_d[n - 1])) continue;
        uint32_t pos = 0;
        for (uint32_t i = lane; i < n; i += 32) pos += res_d[i] <= d ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
        const uint32_t nn = n < p.ef ? n + 1 : n;
        for (int base = (int)nn - 1; base > (int)pos; base -= 32) {      // shift [pos, nn-1) right by one, top chunk first
          const int i = base - (int)lane;
          float vd = 0.0f; uint32_t vs = 0;
          if (i > (int)pos) { vd = res_d[i - 1]; vs = res_s[i - 1]; }
          __syncwarp();
          if (i > (int)pos) { res_d[i] = vd; res_s[i] = vs; }
          __syncwarp();
        }
        if (lane == 0) { res_d[pos] = d; res_s[pos] = s; sh_n = nn; }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  const uint32_t n = sh_n;
  for (uint32_t i = tid; i < n; i += blockDim.x) {
    p.out_slots[(size_t)q * p.ef + i] = res_s[i] & ~kPqExpanded;
    p.out_d2[(size_t)q * p.ef + i] = res_d[i];
  }
  if (tid == 0) {
    p.out_counts[q] = (int)n;
    atomicAdd(p.stats + 0, evals);
    atomicAdd(p.stats + 1, exps);
  }
}

// Final step.  One CTA per query: either re-score the survivors with the reference's exact fp32 arithmetic (2 lanes per
// row x 4 AVX-lane chains, pkg/distance/simd/cpp/avx.cpp:15-32,51-75) and keep the k closest (ties by id), or turn the
// k smallest ADC values into distances (cosine: ||q - r||^2 / 2 for unit vectors; L2: sqrt).
template <int METRIC>
__global__ void __launch_bounds__(kPqThreads) pq_finish_kernel(const uint8_t* rows, uint32_t row_stride, uint32_t dim, uint32_t q_stride, const float* norm2,
                                                               const uint64_t* ids, const float* queries, const float* q_norm2, const uint32_t* slots,
                                                               const float* d2, const int* counts, uint32_t ef, uint32_t k, int rerank, Hit* out,
                                                               int* out_counts) {
  extern __shared__ __align__(16) float fsm[];
  float* q_s = fsm;                                    // [q_stride]
  float* sc = q_s + q_stride;                          // [ef]
  uint64_t* id_s = reinterpret_cast<uint64_t*>(sc + ((ef + 1) & ~1u));   // [ef]
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)counts[q];
  const uint32_t* sl = slots + (size_t)q * ef;
  for (uint32_t d = tid; d < q_stride; d += blockDim.x) q_s[d] = queries[(size_t)q * q_stride + d];
  __syncthreads();
  if (rerank) {
    const float qn = METRIC == COLTT_COSINE ? q_norm2[q] : 0.0f;
    const uint32_t r = lane_row16(lane), g = lane_half(lane);
    const uint32_t full8 = (dim / 8) * 8;
    for (uint32_t base = warp * 16; base < n; base += (blockDim.x >> 5) * 16) {
      const uint32_t j = base + r;
      const bool valid = j < n;
      const uint32_t s = valid ? sl[j] : sl[0];
      const uint8_t* rowp = rows + (size_t)s * row_stride;
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
          float d0 = sub_rn(qv.x, rv[0]), d1 = sub_rn(qv.y, rv[1]), d2_ = sub_rn(qv.z, rv[2]), d3 = sub_rn(qv.w, rv[3]);
          acc[0] = add_rn(acc[0], mul_rn(d0, d0)); acc[1] = add_rn(acc[1], mul_rn(d1, d1));
          acc[2] = add_rn(acc[2], mul_rn(d2_, d2_)); acc[3] = add_rn(acc[3], mul_rn(d3, d3));
        }
      }
      float h = add_rn(add_rn(acc[0], acc[1]), add_rn(acc[2], acc[3]));
      float o = __shfl_xor_sync(0xffffffffu, h, 8);
      float tot = g == 0 ? add_rn(h, o) : add_rn(o, h);
      for (uint32_t d = full8; d < dim; d++) {
        const float rv = reinterpret_cast<const float*>(rowp)[d], qv = q_s[d];
        if (METRIC == COLTT_COSINE) tot = add_rn(tot, mul_rn(qv, rv));
        else { const float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
      }
      if (valid && g == 0) sc[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, norm2[s]) : sqrt_via_f64(tot);
    }
  } else {
    for (uint32_t j = tid; j < n; j += blockDim.x) {
      const float v = d2[(size_t)q * ef + j];
      sc[j] = METRIC == COLTT_COSINE ? fmaxf(0.0f, 0.5f * v) : sqrtf(fmaxf(0.0f, v));
    }
  }
  for (uint32_t j = tid; j < n; j += blockDim.x) id_s[j] = ids[sl[j]];
  __syncthreads();
  const uint32_t n_out = n < k ? n : k;
  for (uint32_t e = tid; e < n; e += blockDim.x) {
    const float se = sc[e];
    const uint64_t ie = id_s[e];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) rank += t_less(sc[j], id_s[j], se, ie) ? 1u : 0u;
    if (rank < n_out) { Hit hh; hh.id = ie; hh.score = se; hh.slot = sl[e]; out[(size_t)q * k + rank] = hh; }
  }
  if (tid == 0) out_counts[q] = (int)n_out;
}

// ---- host ---------------------------------------------------------------------------------------------------------------
int hnsw_pq_train(Hnsw* h, int num_centroids, int num_sub_vectors, int train_rows, int iters) {
  if (num_centroids < 2 || num_centroids > 256) return fail(COLTT_ERR_INVALID, "numCentroids must be in [2, 256]");   // hnsw_common.go:25
  if (num_sub_vectors < 2 || (uint32_t)num_sub_vectors > h->dim || h->dim % (uint32_t)num_sub_vectors)
    return fail(COLTT_ERR_INVALID, "numSubVectors must be >= 2 and divide the dimension");
  const uint32_t M = (uint32_t)num_sub_vectors, C = (uint32_t)num_centroids, dsub = h->dim / M;
  if (dsub > kPqMaxDsub) return fail(COLTT_ERR_UNSUPPORTED, "sub-vectors longer than 32 elements");
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->n < C) return fail(COLTT_ERR_INVALID, "fewer vertices than centroids");
  COLTT_CUDA(cudaSetDevice(h->device));
  const uint32_t T = (uint32_t)std::min<size_t>(h->n, (size_t)std::max(train_rows, num_centroids));
  int rc;
  DeviceBuf sums, counts;
  const size_t cn = (size_t)M * C * dsub;
  if ((rc = h->pq_cent.ensure(cn * 4)) || (rc = h->pq_codes.ensure((size_t)h->n * M)) || (rc = sums.ensure(cn * 4)) || (rc = counts.ensure((size_t)M * C * 4))) return rc;
  cudaStream_t st = h->stream;
  pq_init_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, st>>>(h->d_rows, h->row_stride, h->n, T, M, C, dsub, (float*)h->pq_cent.p);
  count_launch();
  for (int it = 0; it < iters; it++) {
    COLTT_CUDA(cudaMemsetAsync(sums.p, 0, cn * 4, st));
    COLTT_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)M * C * 4, st));
    pq_assign_kernel<true><<<dim3((T + 255) / 256, M), 256, 0, st>>>(h->d_rows, h->row_stride, h->n, T, M, C, dsub, (const float*)h->pq_cent.p, nullptr,
                                                                    (float*)sums.p, (uint32_t*)counts.p);
    pq_update_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, st>>>(M, C, dsub, (const float*)sums.p, (const uint32_t*)counts.p, (float*)h->pq_cent.p);
    count_launch(2);
  }
  pq_assign_kernel<false><<<dim3((unsigned)((h->n + 255) / 256), M), 256, 0, st>>>(h->d_rows, h->row_stride, h->n, T, M, C, dsub, (const float*)h->pq_cent.p,
                                                                                   (uint8_t*)h->pq_codes.p, nullptr, nullptr);
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  COLTT_CUDA(cudaStreamSynchronize(st));
  h->pq_m = M; h->pq_c = C; h->pq_dsub = dsub;
  return COLTT_OK;
}

int hnsw_pq_search(Hnsw* h, const float* queries, size_t nq, int k, int ef_in, int rerank, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(COLTT_ERR_INVALID, "null argument");
  if (k <= 0) return fail(COLTT_ERR_INVALID, "k must be positive");
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->pq_m) return fail(COLTT_ERR_INVALID, "the index has no product quantizer: call coltt_b200_hnsw_pq_train first");
  COLTT_CUDA(cudaSetDevice(h->device));
  if (h->n == 0) { for (size_t q = 0; q < nq; q++) out_counts[q] = 0; return COLTT_OK; }
  const uint32_t ef = (uint32_t)std::max(ef_in > 0 ? ef_in : h->ef_default, k);
  const uint32_t q_stride = (h->dim + 7) / 8 * 8, words = (h->n + 31) / 32;
  const size_t smem = ((size_t)h->pq_m * h->pq_c + 2 * (size_t)ef + 64) * 4;
  if (smem > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "PQ table + ef do not fit shared memory");
  const size_t fsmem = ((size_t)q_stride + ((ef + 1) & ~1u)) * 4 + (size_t)ef * 8;
  cudaStream_t st = h->stream;
  int rc;
  if ((rc = h->q_in.ensure(nq * h->dim * 4)) || (rc = h->q_deq.ensure(nq * q_stride * 4)) || (rc = h->q_n2.ensure(nq * 4)) ||
      (rc = h->visited.ensure(nq * (size_t)words * 4)) || (rc = h->out.ensure(nq * (size_t)k * sizeof(Hit))) || (rc = h->counts.ensure(nq * 4)) ||
      (rc = h->pq_slots.ensure(nq * (size_t)ef * 4)) || (rc = h->pq_d2.ensure(nq * (size_t)ef * 4)) || (rc = h->pq_cnt.ensure(nq * 4)))
    return rc;
  COLTT_CUDA(cudaMemcpyAsync(h->q_in.p, queries, nq * h->dim * 4, cudaMemcpyHostToDevice, st));
  PrepParams pp{};
  pp.in = (const float*)h->q_in.p; pp.n = nq; pp.in_stride = h->dim; pp.dim = h->dim; pp.smem_stride = (h->dim + 3) / 4 * 4;
  pp.normalize = h->metric == COLTT_COSINE;
  pp.norm2_out = (float*)h->q_n2.p; pp.deq_out = (float*)h->q_deq.p; pp.deq_stride = q_stride;
  if ((rc = launch_prep_rows(pp, ELEM_F32, st))) return rc;
  COLTT_CUDA(cudaMemsetAsync(h->visited.p, 0, nq * (size_t)words * 4, st));
  COLTT_CUDA(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(unsigned long long), st));
  PqSearchParams p{};
  p.codes = (const uint8_t*)h->pq_codes.p; p.cent = (const float*)h->pq_cent.p; p.M = h->pq_m; p.C = h->pq_c; p.dsub = h->pq_dsub;
  p.queries = (const float*)h->q_deq.p; p.q_stride = q_stride;
  p.level = h->d_level; p.vbase = h->d_vbase; p.edge_off = h->d_edge_off; p.edge_nbr = h->d_edge_nbr; p.nbr0 = h->d_nbr0; p.nbr0_stride = h->nbr0_stride;
  p.n = h->n; p.entry = h->entry; p.nq = (uint32_t)nq; p.ef = ef;
  p.visited = (uint32_t*)h->visited.p; p.visited_words = words;
  p.out_slots = (uint32_t*)h->pq_slots.p; p.out_d2 = (float*)h->pq_d2.p; p.out_counts = (int*)h->pq_cnt.p; p.stats = h->d_stats;
  if (!h->ev0) { COLTT_CUDA(cudaEventCreate(&h->ev0)); COLTT_CUDA(cudaEventCreate(&h->ev1)); }
  if ((rc = kernel_attrs(pq_search_kernel, smem))) return rc;
  COLTT_CUDA(cudaEventRecord(h->ev0, st));
  pq_search_kernel<<<(unsigned)nq, kPqThreads, smem, st>>>(p);
  count_launch();
  if (h->metric == COLTT_COSINE) {
    if ((rc = kernel_attrs(pq_finish_kernel<COLTT_COSINE>, fsmem))) return rc;
    pq_finish_kernel<COLTT_COSINE><<<(unsigned)nq, kPqThreads, fsmem, st>>>(h->d_rows, h->row_stride, h->dim, q_stride, h->d_norm2, h->d_ids, (const float*)h->q_deq.p,
                                                                           (const float*)h->q_n2.p, (const uint32_t*)h->pq_slots.p, (const float*)h->pq_d2.p,
                                                                           (const int*)h->pq_cnt.p, ef, (uint32_t)k, rerank, (Hit*)h->out.p, (int*)h->counts.p);
  } else {
    if ((rc = kernel_attrs(pq_finish_kernel<COLTT_EUCLIDEAN>, fsmem))) return rc;
    pq_finish_kernel<COLTT_EUCLIDEAN><<<(unsigned)nq, kPqThreads, fsmem, st>>>(h->d_rows, h->row_stride, h->dim, q_stride, h->d_norm2, h->d_ids, (const float*)h->q_deq.p,
                                                                              (const float*)h->q_n2.p, (const uint32_t*)h->pq_slots.p, (const float*)h->pq_d2.p,
                                                                              (const int*)h->pq_cnt.p, ef, (uint32_t)k, rerank, (Hit*)h->out.p, (int*)h->counts.p);
  }
  count_launch();
  COLTT_CUDA(cudaEventRecord(h->ev1, st));
  COLTT_CUDA(cudaGetLastError());
  std::vector<Hit> hits(nq * (size_t)k);
  unsigned long long stats[8];
  COLTT_CUDA(cudaMemcpyAsync(hits.data(), h->out.p, hits.size() * sizeof(Hit), cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaMemcpyAsync(out_counts, h->counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaMemcpyAsync(stats, h->d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->last_kernel_ms = ms;
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

COLTT_API int coltt_b200_hnsw_pq_train(coltt_hnsw* h, const coltt_pq_params* p, int iterations) {
  if (!h || !p) return fail(COLTT_ERR_INVALID, "null argument");
  return coltt::hnsw_pq_train(reinterpret_cast<Hnsw*>(h), p->num_centroids, p->num_sub_vectors, p->trigger_threshold, iterations > 0 ? iterations : 12);
}
COLTT_API int coltt_b200_hnsw_pq_search(coltt_hnsw* h, const float* queries, size_t nq, int k, int ef, int rerank, uint64_t* out_ids, float* out_scores,
                                        int32_t* out_counts) {
  if (!h) return fail(COLTT_ERR_INVALID, "null index");
  return coltt::hnsw_pq_search(reinterpret_cast<Hnsw*>(h), queries, nq, k, ef, rerank, out_ids, out_scores, out_counts);
}

}  // extern "C"
