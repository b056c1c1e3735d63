// kernels.cuh — parameter blocks and launchers of the CUDA kernels (internal header).
#pragma once
#include "common.cuh"

namespace coltt {

// ELEM_F8C: the reference's literal (broken) 8-bit code; ELEM_F8E: builder-defined OCP E4M3 codes + one power-of-two
// scale per row (common.cuh)
enum { ELEM_F32 = 0, ELEM_F16 = 1, ELEM_F8C = 2, ELEM_F8E = 3 };
__host__ __device__ inline uint32_t elem_size(int e) { return e == ELEM_F32 ? 4u : (e == ELEM_F16 ? 2u : 1u); }

// ---- prep.cu ---------------------------------------------------------------------------
struct PrepParams {
  const float* in;        // [n][in_stride] un-normalized fp32 (device)
  size_t n;
  uint32_t in_stride;     // floats
  uint32_t dim;
  uint32_t smem_stride;   // floats per warp scratch (>= dim)
  int normalize;          // cosine
  uint8_t* rows_out;      // nullable: stored rows [slot][row_stride bytes]
  uint32_t row_stride;
  const uint32_t* slots;  // nullable: destination slot per input row
  uint32_t slot_base;     // when slots == nullptr: slot = slot_base + i
  float* norm2_out;       // nullable
  int norm2_by_slot;      // index norm2_out by slot (rows) or by i (queries)
  float* deq_out;         // nullable: fp32 copy of the stored (dequantized) values [n][deq_stride], zero padded
  uint32_t deq_stride;
  __half* f16_out;        // nullable: fp16 copy [n][f16_stride], zero padded
  uint32_t f16_stride;
  uint8_t* shadow_out;    // nullable (ELEM_F32 rows): fp16 copy of the stored row [slot][shadow_stride bytes], zero padded —
  uint32_t shadow_stride; //   the operand of the tensor-core filter for fp32 stores (store.cu)
  float* scale_out;       // ELEM_F8E: per-vector scale, indexed like norm2_out
  uint8_t* code_out;      // nullable (ELEM_F8E queries): the lowered codes [n][code_stride], zero padded
  uint32_t code_stride;
  uint32_t* fill_ptr[3];  // nullable: 32-bit words to initialise in the same launch (search scratch)
  size_t fill_words[3];
  uint32_t fill_value[3];
};
int launch_prep_rows(const PrepParams& p, int elem, cudaStream_t stream);

// ---- flat_scan.cu (K1 exact-order scan + K3 gather front-end) ---------------------------
struct ScanParams {
  const uint8_t* rows;      // [slot][row_stride]
  uint32_t row_stride;      // bytes (multiple of 16)
  uint32_t dim;
  uint32_t n_items;         // rows scanned: n_rows, or n_subset with `subset`
  const uint32_t* subset;   // nullable: slots to gather
  const float* row_norm2;   // [slot]
  const float* row_scale;   // [slot] ELEM_F8E only
  const uint64_t* ids;      // [slot]
  const float* queries;     // [nq][q_stride] dequantized fp32, zero padded to q_stride
  const float* q_norm2;     // [nq]
  uint32_t q_stride;        // floats, multiple of 8
  uint32_t nq;
  uint32_t k;
  int nearest;              // select mode
  int metric;
  Hit* warp_lists;          // scratch [gridDim.x*W][nq][k] best-first
  Hit* cta_lists;           // out     [gridDim.x][nq][k]  best-first
  int* cta_counts;          // out     [gridDim.x][nq]
  const uint32_t* q_map;    // nullable: compact query index -> prepared query row
  const uint32_t* n_active; // nullable: number of queries, on the device (overrides nq)
  uint32_t q_base;          // with n_active: this launch serves compact indices [q_base, q_base + nq)
  uint32_t chunk_bytes;     // bytes of one row copied per pipeline stage (multiple of 128)
  uint32_t n_stages;        // stages per warp
  // CFLAT multi-vector scoring (experimental/multi_vector_vertex.go:108-116): one launch per included field adds
  // scoreHelper(distance) * weight to a per-slot running score; the launch of the last field selects on the sum.
  float* multi_acc;         // nullable: [slot] running score (nq must be 1)
  float multi_w;            // float32(ratio) / 100
  int multi_first, multi_last;
};
struct ScanPlan {
  int grid_x, grid_y, warps, qt;
  uint32_t chunk_bytes, n_stages;
  size_t smem_bytes;
  size_t warp_list_bytes, cta_list_bytes, cta_count_bytes;
};
// Chooses tile/pipeline shape for (elem, dim, nq, k, n_items); no device work.
int plan_flat_scan(int elem, uint32_t dim, uint32_t row_stride, uint32_t n_items, uint32_t nq, uint32_t k, int n_sms,
                   ScanPlan* plan);
int launch_flat_scan(const ScanParams& p, const ScanPlan& plan, int elem, cudaStream_t stream);

// ---- gemm_filter.cu (K2) + rerank.cu ------------------------------------------------------
struct GemmCand {
  float key;      // "larger is better" approximate score (see gemm_filter.cu)
  uint32_t row;   // row slot
};
struct GemmParams {
  uint32_t n_rows, dim, nq;
  const void* q_lowered;     // [nq][q_stride bytes] lowered queries (fp16, or E4M3 codes), zero padded to kblocks*128 bytes
  uint32_t q_stride;         // bytes
  const float* row_norm2;    // [n_rows]
  const float* row_scale;    // [n_rows] E4M3 stores only (nullable): folded into the per-row key coefficient
  const uint8_t* rows;       // the shard's row matrix (the operands themselves go through TMA)
  uint32_t row_stride;       // bytes
  int metric, nearest;
  uint32_t* g_thr;           // [nq] order-encoded bound the survivors were cut at (max over columns), zero-initialised
  float* pub;                // [n_cols][nq] the order statistic each column publishes per query, initialised to NaN
  GemmCand* cand_buf;        // scratch [grid_y*grid_x*2][cand_cap][128]: per-(column,query) candidate buffers
  GemmCand* cand_out;        // [nq][n_cols][cand_out_cap]
  uint32_t* cand_cnt;        // [nq][n_cols]
  uint32_t kblocks, kprime, cand_cap, cand_out_cap, n_stages;  // filled from the plan
  uint32_t pub_kth, groups;  // bound scheme (gemm_common.cuh): what a column publishes, and the number of column classes
  float* dbg_acc;            // nullable (tests): raw accumulators [nq][n_rows]
  unsigned long long* dbg_prof;  // nullable (profiling builds, -DCOLTT_K2_PROF=1): [grid][8] cycle counters per role
  unsigned long long* dbg_prof2; // second bank of counters (epilogue detail)
  uint32_t pf_inner;         // L2 prefetch box width in bytes (multiple of 64)
  uint32_t tma_shift;        // log2 of the tensor maps' word width in bytes: TMA coordinates = byte offsets >> tma_shift
  uint32_t dbg_flags;        // profiling builds only. bit0: epilogue only drains TMEM; bit1: no L2 prefetch
};
struct GemmPlan {
  uint32_t kblocks, kprime, kp, pub_kth, groups, cand_cap, cand_out_cap, n_stages, grid_x, grid_y, q_stride, tile_rows, pair, n_cols, fp8;
  uint32_t sb;               // bytes per row of a shard-tile stage (64 or 128; CTA-pair kernel)
  size_t smem_bytes;
};
// The three TMA descriptors of a launch (shard tiles, query tile, L2-prefetch view), kept with the search scratch
// and re-encoded only when the tensor they describe changes (cuTensorMapEncodeTiled is host work per search).
struct GemmMapCache {
  alignas(64) unsigned char maps[3][128];   // CUtensorMap x3 (opaque here: this header does not pull in <cuda.h>)
  const void* rows = nullptr;
  const void* q = nullptr;
  uint32_t n_rows = 0, row_bytes = 0, row_stride = 0, box_rows = 0, pf_inner = 0, nq = 0, q_stride = 0, sb = 0, ew = 0, q_ew = 0;
};
// row_bytes = dim * element size; fp8 = E4M3 operands (kind::f8f6f4) instead of fp16 (kind::f16)
int plan_gemm_filter(uint32_t row_bytes, bool fp8, uint32_t nq, uint32_t k, int n_sms, GemmPlan* plan);
int launch_gemm_filter(const GemmParams& p, const GemmPlan& plan, const void* d_rows, uint32_t row_stride, cudaStream_t stream,
                       GemmMapCache* cache = nullptr);
uint32_t gemm_filter_cols(const GemmPlan& plan, uint32_t n_rows);

struct RerankParams {
  uint32_t nq, k, dim, q_stride, row_stride, grid_x, cand_cap;
  uint32_t max_rows;         // rows re-scored exactly per query: 64 (top-10/24) or 256 (top-100)
  uint32_t chunk_rows;       // rows staged in shared memory at a time (filled by launch_rerank)
  float eps_rel;             // certificate margin: |filter score - exact score| <= eps_rel * ||q|| ||row|| (fast_eps_rel(dim))
  int metric, nearest, elem;
  const float* queries;      // [nq][q_stride] dequantized fp32 (exact path operand)
  const float* q_norm2;      // [nq]
  const float* q_scale;      // [nq] E4M3 stores only (nullable): the filter's keys are in units of 1/q_scale
  const uint8_t* rows;
  const float* row_norm2;
  const float* row_scale;    // [slot] E4M3 stores only
  const uint64_t* ids;
  const GemmCand* cand_in;   // [nq][grid_x][cand_cap]
  const uint32_t* cand_cnt;  // [nq][grid_x]
  const uint32_t* g_thr;     // [nq]
  Hit* out;                  // [nq][k_out_stride] in T order
  uint32_t out_stride;
  int* out_counts;           // [nq]
  uint32_t* flags;           // [nq] 1 = margin not certified: the caller re-runs that query EXACT
  // The last CTA to finish compacts the flagged queries for the exact re-run that follows in the stream:
  uint32_t* done_ctr;        // arrival counter, zero before the launch; the last CTA resets it
  uint32_t* q_map;           // [nq] out: the flagged queries
  uint32_t* n_bad;           // [1]  out: how many (always written)
  unsigned long long* stat_fallbacks;  // nullable: running total of flagged queries (statistics)
};
int launch_rerank(const RerankParams& p, cudaStream_t stream);
// Certificate margin of COLTT_MATH_FAST, relative to ||q|| ||row|| (DESIGN.md §5 has the derivation): the tensor core adds
// `dim` exact products into an fp32 accumulator truncating each addend to the accumulator's ulp (<= dim * 2^-23), the exact
// kernel rounds (dim/8 + 3) times per AVX lane (<= (dim/8+3) * 2^-24); 25 % slack and 2^-20 for the epilogue roundings.
// fp32 stores are filtered through an fp16 shadow of the (unit-norm) rows and query: every element is off by at most 2^-11
// relative (round to nearest; 2^-25 absolute below 2^-14), so the dot product moves by <= (2^-10 + 2^-22) ||q|| ||row|| plus
// 2^-24 sqrt(dim); added to the accumulation-order margin above.
inline float fast_eps_rel_f32_shadow(uint32_t dim);
inline float fast_eps_rel(uint32_t dim) {
  return 1.25f * ((float)dim * 1.1920929e-7f + ((float)dim / 8.0f + 3.0f) * 5.9604645e-8f) + 9.5367432e-7f;
}
inline float fast_eps_rel_f32_shadow(uint32_t dim) {
  float r = 1.0f;
  while (r * r < (float)dim) r += 1.0f;      // ceil(sqrt(dim)) without <cmath> in this header
  return fast_eps_rel(dim) + 9.7656250e-4f + 2.3841858e-7f + 5.9604645e-8f * r;
}

// ---- topk_merge.cu (K5) -----------------------------------------------------------------
struct MergeParams {
  const Hit* lists;          // [n_lists][nq][k_in]
  const int* counts;         // [n_lists][nq]
  int n_lists;
  uint32_t nq;
  uint32_t k_in;
  uint32_t k;
  int nearest;
  int in_best_first;         // 1: lists are best-first (internal); 0: lists are in T order (public)
  const uint32_t* q_map;     // nullable: compact query index -> output row
  const uint32_t* n_active;  // nullable: number of queries, on the device
  uint32_t q_base;           // with n_active: this launch serves compact indices [q_base, q_base + nq)
  uint32_t out_stride;       // hits per output row (0 = k)
  size_t list_stride_hits;   // Hits between consecutive lists (0 = nq*k_in, the dense layout)
  size_t count_stride;       // ints between consecutive lists' counts (0 = nq)
  // Peer-memory form (comm.cu): list j lives at its own base address — another GPU's exchange buffer, read over NVLink —
  // as coltt_hit[nq][k_in] followed, counts_off bytes in, by int32 counts[nq].  Null = the lists / counts arrays above.
  const unsigned long long* list_bases;   // device array [n_lists]
  size_t counts_off;
  Hit* out;                  // [nq][out_stride] in T order (ascending score, NaN last, then id)
  int* out_counts;           // [nq]
};
int launch_merge_topk(const MergeParams& p, cudaStream_t stream);

}  // namespace coltt
