// rerank.cu — exact re-scoring of the tcgen05 filter's survivors + margin certificate.
//
// Second half of COLTT_MATH_FAST.  For one query per CTA: gather the candidates every filter CTA
// kept (gemm_filter.cu), drop those below the final shared threshold, re-score the rest with
// the EXACT arithmetic of flat_scan.cu (the reference's AVX evaluation order,
// pkg/distance/simd/cpp/avx.cpp:51-75 -> simd/avx/AVX_amd64.go:44-52 -> space.go:93-95), select
// the top K in T order and certify that no row the filter dropped can beat the K-th:
//   every dropped row has approximate key <= B (the final threshold), the approximate and exact
//   scores differ by at most eps (fp32 accumulation-order error over exact fp16 x fp16 products),
//   so the answer is provably the EXACT answer when the exact K-th score clears B by eps.
// Queries that cannot be certified (heavy ties at the threshold) are flagged and re-run on the
// exact path by the caller (the last CTA to finish compacts them into a list).  Traffic: ~2K rows of dim*2 bytes per query — negligible next to the scan.
#include "exact_math.cuh"
#include "store.h"
#include "topk.cuh"

namespace coltt {

static constexpr int kRerankThreads = 256;
static constexpr uint32_t kRerankChunkMax = 64;    // rows staged in shared memory at a time (fewer when rows are wide)

// shared-memory layout (byte offsets from the dynamic base; offsets, not rounded pointers, so that every access
// stays in the shared address space and compiles to LDS/STS).  MC = survivors gathered per query (keys only),
// MR = rows re-scored exactly per query (>= 2 K').
struct RerankSmem {
  uint32_t q, row, key, sel, score, n2, sc, id, rows, total;
  __host__ __device__ RerankSmem(uint32_t q_stride, uint32_t row_stride, uint32_t MC, uint32_t MR, uint32_t chunk) {
    q = 0;                                   // float  [q_stride]        (bulk-copied: 16-byte multiple)
    row = q + q_stride * 4;                  // u32    [MC]
    key = row + MC * 4;                      // float  [MC]
    sel = key + MC * 4;                      // u32    [MR]
    score = sel + MR * 4;                    // float  [MR]
    n2 = score + MR * 4;                     // float  [MR]     ||row||^2 of the selected rows
    sc = n2 + MR * 4;                        // float  [MR]     E4M3 row scales
    id = sc + MR * 4;                        // u64    [MR]
    rows = (id + MR * 8 + 127u) & ~127u;     // bytes [chunk][row_stride + 16]
    total = rows + chunk * (row_stride + 16);
  }
};

template <int ELEM, int METRIC, int MR>
__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(RerankParams p) {
  constexpr uint32_t MC = MR == 64 ? 1024u : 4096u;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t n_s, ovf_s, last_s;
  __shared__ float kth_s, bound_s;
  __shared__ __align__(8) uint64_t bar_s[2];   // [0] query copy, [1] row copies (one phase per chunk)
  const RerankSmem L(p.q_stride, p.row_stride, MC, MR, p.chunk_rows);
  const uint32_t kRerankChunk = p.chunk_rows;
  float* q_s = reinterpret_cast<float*>(smem + L.q);
  uint32_t* row_s = reinterpret_cast<uint32_t*>(smem + L.row);
  float* key_s = reinterpret_cast<float*>(smem + L.key);
  uint32_t* sel_row = reinterpret_cast<uint32_t*>(smem + L.sel);
  float* score_s = reinterpret_cast<float*>(smem + L.score);
  float* n2_s = reinterpret_cast<float*>(smem + L.n2);
  float* sc_s = reinterpret_cast<float*>(smem + L.sc);
  uint64_t* id_s = reinterpret_cast<uint64_t*>(smem + L.id);
  uint8_t* rows_s = smem + L.rows;
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t ES = ELEM == ELEM_F32 ? 4 : (ELEM == ELEM_F16 ? 2 : 1);
  const uint32_t RS = p.row_stride + 16;   // padded shared-memory row stride (bank-conflict free, as in flat_scan.cu)

  if (tid == 0) {
    n_s = 0; ovf_s = 0; kth_s = 0.0f;
    mbar_init(smem_u32(&bar_s[0]), 1);
    mbar_init(smem_u32(&bar_s[1]), 1);
    fence_mbar_init();
    // the query streams into shared memory while the survivors are gathered and ranked
    mbar_arrive_expect_tx(smem_u32(&bar_s[0]), p.q_stride * 4);
    bulk_g2s(smem_u32(q_s), p.queries + (size_t)q * p.q_stride, p.q_stride * 4, smem_u32(&bar_s[0]));
  }
  // padding keys: never better than anything, never tie-break ahead of anything
  for (uint32_t e = tid; e < MC; e += blockDim.x) { key_s[e] = __int_as_float(0xff800000); row_s[e] = 0xffffffffu; }
  __syncthreads();

  // ---- 1. gather the survivors of every filter column that clear the final threshold.  One column per
  //         thread, two columns in flight: count, then its few entries four at a time
  const uint32_t thr_bits = p.g_thr[q];
  const bool have_bound = thr_bits != 0;
  const float B = have_bound ? ord2f(thr_bits) : 0.0f;
  for (uint32_t c0 = tid; c0 < p.grid_x; c0 += 2 * blockDim.x) {
    uint32_t ccnt[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t cta = c0 + u * blockDim.x;
      ccnt[u] = cta < p.grid_x ? p.cand_cnt[(size_t)q * p.grid_x + cta] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t cta = c0 + u * blockDim.x;
      if (ccnt[u] == 0xffffffffu) { ovf_s = 1; continue; }           // that column overflowed on ties: exact path
      const uint32_t cn = ccnt[u] < p.cand_cap ? ccnt[u] : p.cand_cap;
      const GemmCand* src = p.cand_in + ((size_t)q * p.grid_x + cta) * p.cand_cap;
      for (uint32_t s0 = 0; s0 < cn; s0 += 4) {
        GemmCand e4[4];
#pragma unroll
        for (int v = 0; v < 4; v++) if (s0 + v < cn) e4[v] = src[s0 + v];
#pragma unroll
        for (int v = 0; v < 4; v++) {
          if (s0 + v < cn && (!have_bound || e4[v].key >= B || e4[v].key != e4[v].key)) {
            const uint32_t pos = atomicAdd(&n_s, 1u);
            if (pos < MC) { row_s[pos] = e4[v].row; key_s[pos] = e4[v].key; }
          }
        }
      }
    }
  }
  __syncthreads();
  const uint32_t n_all = n_s;
  const uint32_t n = n_all < MC ? n_all : MC;

  // ---- 2. keep the M best by approximate key (rank counting; ties by row) — everything else, gathered or
  //         not, has key <= bound': the (M+1)-th best key, or B when fewer than M were gathered
  const uint32_t M = n < (uint32_t)MR ? n : (uint32_t)MR;
  if (tid == 0) bound_s = have_bound ? B : __int_as_float(0xff800000);
  __syncthreads();
  const uint32_t n4 = (n + 3) & ~3u;         // the padding entries rank behind everything
  for (uint32_t e = tid; e < n; e += blockDim.x) {
    const float ke = key_s[e];
    const uint32_t re = row_s[e];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n4; j += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(key_s + j);
      rank += (k4.x > ke ? 1u : 0u) + (k4.y > ke ? 1u : 0u) + (k4.z > ke ? 1u : 0u) + (k4.w > ke ? 1u : 0u);
      if (k4.x == ke || k4.y == ke || k4.z == ke || k4.w == ke) {   // ties are rare: the row order decides
        const uint4 r4 = *reinterpret_cast<const uint4*>(row_s + j);
        rank += (k4.x == ke && r4.x < re ? 1u : 0u) + (k4.y == ke && r4.y < re ? 1u : 0u) + (k4.z == ke && r4.z < re ? 1u : 0u) +
                (k4.w == ke && r4.w < re ? 1u : 0u);
      }
    }
    if (rank < M) sel_row[rank] = re;
    if (rank == M && n > M) bound_s = fmaxf(bound_s, ke);   // single writer: ranks are unique
  }
  __syncthreads();

  // ---- 3. fetch the M rows, kRerankChunk at a time, with one bulk async copy each (a chunk's copies are all in
  //         flight at once; norms, scales and ids are fetched by the same threads meanwhile), and re-score them from
  //         shared memory with the exact AVX-order arithmetic of flat_scan.cu
  mbar_wait(smem_u32(&bar_s[0]), 0);
  const uint32_t r = lane_row16(lane), g = lane_half(lane);
  const uint32_t full8 = (p.dim / 8) * 8;
  const float qn = METRIC == COLTT_COSINE ? p.q_norm2[q] : 0.0f;
  for (uint32_t j = tid; j < M; j += blockDim.x) {
    const uint32_t row = sel_row[j];
    n2_s[j] = METRIC == COLTT_COSINE ? p.row_norm2[row] : 0.0f;
    sc_s[j] = ELEM == ELEM_F8E ? p.row_scale[row] : 1.0f;
    id_s[j] = p.ids[row];
  }
  for (uint32_t c0 = 0, ph = 0; c0 < M; c0 += kRerankChunk, ph ^= 1) {
    const uint32_t mc = M - c0 < kRerankChunk ? M - c0 : kRerankChunk;
    if (tid == 0) mbar_arrive_expect_tx(smem_u32(&bar_s[1]), mc * p.row_stride);
    __syncthreads();    // the previous chunk's reads are done; expect_tx precedes the copies; n2_s / sc_s / id_s are in place
    for (uint32_t j = tid; j < mc; j += blockDim.x)
      bulk_g2s(smem_u32(rows_s + (size_t)j * RS), p.rows + (size_t)sel_row[c0 + j] * p.row_stride, p.row_stride, smem_u32(&bar_s[1]));
    mbar_wait(smem_u32(&bar_s[1]), ph);
    for (uint32_t base = warp * 16; base < mc; base += (blockDim.x >> 5) * 16) {
      const uint32_t jl = base + r;
      const bool valid = jl < mc;
      const uint32_t j = c0 + (valid ? jl : 0);
      const uint8_t* rowp = rows_s + (size_t)(valid ? jl : 0) * RS;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (uint32_t e = 0; e < full8; e += 8) {
        float rv[4];
        load4<ELEM>(rowp + (size_t)(e + 4 * g) * ES, nullptr, rv);
        const float4 qv = *reinterpret_cast<const float4*>(q_s + e + 4 * g);
        if (METRIC == COLTT_COSINE) {
          acc[0] = dot_step<ELEM>(acc[0], qv.x, rv[0]); acc[1] = dot_step<ELEM>(acc[1], qv.y, rv[1]);
          acc[2] = dot_step<ELEM>(acc[2], qv.z, rv[2]); acc[3] = dot_step<ELEM>(acc[3], qv.w, rv[3]);
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
        float rv = load1<ELEM>(rowp, d, nullptr);
        float qv = q_s[d];
        if (METRIC == COLTT_COSINE) tot = dot_step<ELEM>(tot, qv, rv);
        else { float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
      }
      if (ELEM == ELEM_F8E) tot = mul_rn(tot, sc_s[j]);   // the row's power-of-two scale, applied once (flat_scan.cu)
      if (valid && g == 0) score_s[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, n2_s[j]) : sqrt_via_f64(tot);
    }
  }
  __syncthreads();

  // ---- 4. top-K of the exact scores, straight into T order
  const uint32_t n_out = M < p.k ? M : p.k;
  Hit* out = p.out + (size_t)q * p.out_stride;
  for (uint32_t e = tid; e < M; e += blockDim.x) {
    const float sc = score_s[e];
    const uint64_t id = id_s[e];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < M; j++) rank += better(score_s[j], id_s[j], sc, id, p.nearest) ? 1u : 0u;
    if (rank < n_out) {
      Hit hh; hh.id = id; hh.score = sc; hh.slot = sel_row[e];
      out[p.nearest ? rank : n_out - 1 - rank] = hh;
      if (rank == n_out - 1) kth_s = sc;
    }
  }
  __syncthreads();

  // ---- 5. certificate: every row that was not re-scored has approximate key <= bound'
  if (tid == 0) {
    p.out_counts[q] = (int)n_out;
    bool ok = n_all <= MC && ovf_s == 0;
    const float Bp = bound_s;
    if (ok && Bp > __int_as_float(0xff800000)) {
      const float dK = kth_s;  // exact score (a distance) of the worst row we return
      if (!(dK == dK)) {
        ok = !p.nearest;       // NaN is the best COMPAT score and the worst NEAREST one
      } else if (METRIC == COLTT_COSINE) {
        // key = +-dot' * s_row/||row|| with dot' in units of 1/s_q (E4M3) — the exact similarity lies within
        // eps_rel of key * s_q/||q||; distance = |1 - sim|
        const float invq = rsqrtf(qn) * (p.q_scale ? p.q_scale[q] : 1.0f);
        const float eps = p.eps_rel;
        if (p.nearest) ok = dK < 1.0f - Bp * invq - eps;        // dropped rows: distance >= 1 - B'*invq - eps
        else ok = dK > 1.0f + Bp * invq + eps;                  // dropped rows: distance <= 1 + B'*invq + eps
      } else {
        // key = +-(2 dot - ||row||^2); the dot product is off by at most eps_rel * ||q|| ||row||, and
        // 2 ||q|| ||row|| <= ||q||^2 + (||q|| + d)^2 for a row at distance d
        const float nq2 = p.q_norm2[q];
        const float d2 = dK * dK;
        const float scale = nq2 + (sqrtf(nq2) + dK) * (sqrtf(nq2) + dK);
        const float eps = p.eps_rel * scale;
        if (p.nearest) ok = d2 < nq2 - Bp - eps;                // key = 2dot - ||row||^2  =>  d^2 = ||q||^2 - key
        else ok = d2 > nq2 + Bp + eps;                          // key = ||row||^2 - 2dot  =>  d^2 = ||q||^2 + key
      }
      if (n_out < p.k) ok = false;  // the filter keeps K' >= K rows once it has a bound: fewer means trouble
    }
    p.flags[q] = ok ? 0u : 1u;
    // ---- 6. last CTA out builds the list of uncertified queries (replaces a separate compaction launch,
    //         a memset of its counter and a device->host copy of the statistic)
    __threadfence();
    last_s = atomicAdd(p.done_ctr, 1u) == gridDim.x - 1 ? 1u : 0u;
    n_s = 0;
  }
  __syncthreads();
  if (last_s) {
    __threadfence();
    for (uint32_t j = tid; j < p.nq; j += blockDim.x)
      if (__ldcg(p.flags + j)) p.q_map[atomicAdd(&n_s, 1u)] = j;
    __syncthreads();
    if (tid == 0) {
      *p.n_bad = n_s;
      if (p.stat_fallbacks && n_s) atomicAdd(p.stat_fallbacks, (unsigned long long)n_s);
      *p.done_ctr = 0;      // ready for the next launch on this scratch (stream-ordered)
    }
  }
}

int launch_rerank(const RerankParams& p_in, cudaStream_t stream) {
  if (p_in.nq == 0) return COLTT_OK;
  if (p_in.max_rows != 64 && p_in.max_rows != 256) return fail(COLTT_ERR_INVALID, "rerank: max_rows must be 64 or 256");
  RerankParams p = p_in;
  const uint32_t MCv = p.max_rows == 64 ? 1024u : 4096u;
  p.chunk_rows = kRerankChunkMax;
  while (p.chunk_rows > 16 && RerankSmem(p.q_stride, p.row_stride, MCv, p.max_rows, p.chunk_rows).total > 200 * 1024) p.chunk_rows -= 16;
  const size_t smem = RerankSmem(p.q_stride, p.row_stride, MCv, p.max_rows, p.chunk_rows).total;
  if (smem > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "rerank: rows too wide for shared memory");
  const bool cosine = p.metric == COLTT_COSINE;
#define COLTT_RR(E, M, R)                                                                                 \
  {                                                                                                       \
    auto kfn = rerank_kernel<E, M, R>;                                                                    \
    { int arc = kernel_attrs(kfn, smem); if (arc) return arc; }                                           \
    kfn<<<p.nq, kRerankThreads, smem, stream>>>(p);                                                        \
  }
#define COLTT_RR_R(E, M) { if (p.max_rows == 64) COLTT_RR(E, M, 64) else COLTT_RR(E, M, 256) }
  if (p.elem == ELEM_F16) { if (cosine) COLTT_RR_R(ELEM_F16, COLTT_COSINE) else COLTT_RR_R(ELEM_F16, COLTT_EUCLIDEAN) }
  else if (p.elem == ELEM_F32) { if (cosine) COLTT_RR_R(ELEM_F32, COLTT_COSINE) else COLTT_RR_R(ELEM_F32, COLTT_EUCLIDEAN) }
  else if (p.elem == ELEM_F8E && cosine) COLTT_RR_R(ELEM_F8E, COLTT_COSINE)
  else return fail(COLTT_ERR_UNSUPPORTED, "rerank: element type / metric");
#undef COLTT_RR_R
#undef COLTT_RR
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
