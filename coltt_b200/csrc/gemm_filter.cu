// gemm_filter.cu — K2: batched query x shard distance as a tcgen05 GEMM with a fused
// threshold top-K' epilogue (COLTT_MATH_FAST), for fp16 rows ("bf16"/f16 stores, tcgen05 kind::f16) and
// E4M3 rows (the builder-defined F8_E4M3 store, kind::f8f6f4 — same byte geometry, twice the elements per MMA).
//
// What it replaces: the same VertexSearch hot loop as flat_scan.cu
// (edge/bf16_vectorstore.go:131-186 -> bf16_quantization.go:33-43 -> pkg/distance), for a
// whole batch of queries at once.  The reference lowers the QUERY to fp16 too
// (bf16_vectorstore.go:136), so both operands are fp16 and every product is exact in fp32 —
// exactly tensor-core semantics.  Only the accumulation order differs from the AVX
// path, so this kernel is used as a FILTER: it keeps, per query, every row whose approximate
// score could be in the top K' (K' = 2K); rerank.cu re-scores the survivors with the exact
// AVX-order arithmetic and certifies the margin, so the returned ids/scores equal the EXACT path.
//
// Mapping (one CTA per SM, persistent over shard tiles):
//   M = 128 queries: the A operand, RESIDENT IN SHARED MEMORY for the whole kernel (12 K-blocks of
//                       128 rows x 128 B, 128B-swizzled, loaded once by TMA = 192 KB of the 227 KB).
//                       (v1 kept it in tensor memory; ncu + in-kernel timers showed every TS-mode MMA then
//                       re-reads its 4 KB A slice from TMEM in ~128 cycles, 4x the N=64 math time.)
//   N = 64 shard rows per tile -> B operand, K-major (= the row-major shard as stored), streamed by TMA
//                       (cp.async.bulk.tensor, 128B swizzle) through a 4-stage x 8 KB mbarrier ring; an L2
//                       prefetch (cp.async.bulk.prefetch.tensor) runs 3 tiles ahead so the short ring only
//                       has to cover L2 latency, not HBM latency.
//   D = 128 x 64 fp32 accumulators, double-buffered in 128 TMEM columns: the epilogue of tile i overlaps
//                       the MMAs of tile i+1.
//   Warp 0: TMA producer.  Warp 1: tcgen05.mma issuer (one elected lane, SS descriptors).  Warps 2-5:
//   epilogue — thread = query: tcgen05.ld the 64 scores of its query, one FFMA turns each into a "larger is
//   better" key (cosine: +-dot/||row||; L2: +-(2 dot - ||row||^2)), a max tree and ONE warp vote per 32
//   scores decide whether anything beats the running K'-th bound held in a register; the rare survivors are
//   appended to an L2-resident per-(CTA,query) buffer.  The bound is shared across CTAs: each CTA publishes
//   its best key per query, and K' disjoint groups of CTAs each holding a row >= x make x a valid bound.
// No second pass over HBM: bytes per launch = N*dim*2 (+ N*4 norms + nq*dim*2 queries).
#include <cstdlib>

#include "gemm_common.cuh"
#include "store.h"

namespace coltt {

template <int KP, bool FP8>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_filter_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_pf, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // no static __shared__ in this kernel: the window base is 1024-aligned
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t NS = p.n_stages, KB = p.kblocks;
  constexpr uint32_t STAGE_BYTES = kBN * kBKB;      // 256 rows x 64 B = 16 KB
  constexpr uint32_t ABLK_BYTES = 128 * kBK;        // 128 queries x 128 B = 16 KB per K block
  const uint32_t NSTEP = KB * (kBK / kBKB);         // shard-tile stages per tile (2 per query K block)

  uint8_t* a_smem = smem;                                   // [KB][128 rows][128 B], 128B-swizzled, resident
  uint8_t* b_stages = a_smem + (size_t)KB * ABLK_BYTES;     // [NS][256 rows][64 B], 64B-swizzled
  float* coef_a = reinterpret_cast<float*>(b_stages + (size_t)NS * STAGE_BYTES);  // [kBN]
  float* coef_b = coef_a + kBN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(coef_b + kBN);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;   // [2]
  uint64_t* a_bar = tempty_bar + 2;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();              // swizzled operands need a 1024-byte aligned base

  const uint32_t n_tiles = (p.n_rows + kBN - 1) / kBN;
  const uint32_t q_tile0 = blockIdx.y * 128;
  const uint32_t cta_lin = blockIdx.y * gridDim.x + blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < NS; s++) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(smem_u32(tfull_bar + b), 1); mbar_init(smem_u32(tempty_bar + b), 8); }
    mbar_init(smem_u32(a_bar), 1);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pf) : "memory");
  }
  if (warp == 1) {   // two 128x256 fp32 accumulators = all of tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    // Whole warp walks the loop (warp-uniform values stay in uniform registers), one elected lane issues.  The
    // ring is shorter than the L2 round trip, so "stage free" -> "load issued" is on the critical path: the load
    // goes out first, the L2 prefetch of the next tile (same K slice, one tile ahead) after it.
    if (elect_one()) {
      // the query tile: A operand, loaded once (rows beyond nq are zero-filled by TMA)
      mbar_arrive_expect_tx(smem_u32(a_bar), KB * ABLK_BYTES);
      for (uint32_t kb = 0; kb < KB; kb++)
        tma_load_2d(smem_u32(a_smem + (size_t)kb * ABLK_BYTES), &tmap_q, (int)((kb * kBK) >> p.tma_shift), (int)q_tile0, smem_u32(a_bar));
    }
    __syncwarp();
    uint32_t s = 0, ph = 0;
#if COLTT_K2_PROF
    const bool do_pf = (p.dbg_flags & 2u) == 0;
#else
    constexpr bool do_pf = true;
#endif
    const uint32_t pf_mask = p.pf_inner / kBKB - 1;          // pf_inner / kBKB is a power of two
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(b_stages);
    if (do_pf && blockIdx.x < n_tiles && elect_one())
      for (uint32_t st = 0; st < NSTEP; st += pf_mask + 1) tma_prefetch_2d(&tmap_pf, (int)((st * kBKB) >> p.tma_shift), (int)(blockIdx.x * kBN));
    __syncwarp();
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int row = (int)(t * kBN), row_pf = row + (int)(gridDim.x * kBN);
      const bool pf = do_pf && t + gridDim.x < n_tiles;
      for (uint32_t st = 0; st < NSTEP; st++) {
        mbar_wait(empty0 + s * 8, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full0 + s * 8, STAGE_BYTES);
          tma_load_2d(stage0 + s * STAGE_BYTES, &tmap, (int)((st * kBKB) >> p.tma_shift), row, full0 + s * 8);
          if (pf && (st & pf_mask) == 0) tma_prefetch_2d(&tmap_pf, (int)((st * kBKB) >> p.tma_shift), row_pf);
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the loop (warp-uniform control flow); one elected lane issues tcgen05.mma / commit.
    // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b F16 K-major, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t s = 0, ph = 0, ti = 0;
    long long w_tempty = 0, w_full = 0, t_start = K2_NOW();
    mbar_wait(smem_u32(a_bar), 0);
    tc_fence_after();
    // single-thread critical path: ring position, phase and descriptors advance by constant adds
    const uint64_t a_desc0 = make_desc_sw128(smem_u32(a_smem));
    const uint64_t b_desc0 = make_desc_sw64(smem_u32(b_stages));
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
      const uint32_t buf = ti & 1, bph = (ti >> 1) & 1;
      const long long c0 = K2_NOW();
      mbar_wait(smem_u32(tempty_bar + buf), bph ^ 1);    // epilogue drained this accumulator
      w_tempty += K2_NOW() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kBN;
      uint64_t a_desc = a_desc0;
      for (uint32_t st = 0; st < NSTEP; st++) {
        const long long c1 = K2_NOW();
        mbar_wait(full0 + s * 8, ph);
        w_full += K2_NOW() - c1;
        tc_fence_after();
        const uint64_t b_desc = b_desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
        if (elect_one()) {
          umma_ss<FP8>(d_tmem, a_desc, b_desc, idesc, st != 0 ? 1u : 0u);
          umma_ss<FP8>(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);     // +32 B: the next MMA's K slice
          umma_commit(empty0 + s * 8);                    // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        a_desc += (st & 1) ? (uint64_t)((ABLK_BYTES - 64) >> 4) : 4ull;   // 32 elements = 64 B into the 128 B swizzle row, then the next K block
        if (++s == NS) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(smem_u32(tfull_bar + buf));   // accumulator complete
      __syncwarp();
    }
#if COLTT_K2_PROF
    if (p.dbg_prof && lane == 0) {
      p.dbg_prof[(size_t)cta_lin * 8 + 2] = (unsigned long long)w_tempty;
      p.dbg_prof[(size_t)cta_lin * 8 + 3] = (unsigned long long)w_full;
      p.dbg_prof[(size_t)cta_lin * 8 + 4] = (unsigned long long)(K2_NOW() - t_start);
    }
#else
    (void)w_tempty; (void)w_full; (void)t_start;
#endif
  } else {
    auto arrive = [&](uint32_t buf) { mbar_arrive(smem_u32(tempty_bar + buf)); };
    filter_epilogue<KP>(p, tmem_base, coef_a, coef_b, tfull_bar, arrive, blockIdx.x, gridDim.x, n_tiles, q_tile0, blockIdx.x, gridDim.x, cta_lin, cta_lin);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int launch_gemm_filter_pair(const CUtensorMap& tm, const CUtensorMap& tmq, const CUtensorMap& tmpf, const GemmParams& p, const GemmPlan& plan,
                            cudaStream_t stream);

// Chooses the bound scheme (gemm_common.cuh) and the launch shape for a shard of `row_bytes`-wide rows.
//   top-k <= 10: K' = 16, KP = 16 registers, columns publish their best key;   top-k <= 24: K' = 32, KP = 32;
//   top-k <= 200: KP = 16, columns publish their 16th best key, groups = ceil(1.25 k / 16), K' = 16 groups
//   (top-100: groups 8, K' = 128).  Larger k, or fewer columns than groups, is served by the exact path.
int plan_gemm_filter(uint32_t row_bytes, bool fp8, uint32_t nq, uint32_t k, int n_sms, GemmPlan* plan) {
  const uint32_t kblocks = (row_bytes + kBK - 1) / kBK;
  uint32_t kp, kprime, pub_kth, groups, cap, out_cap;
  if (k <= 10) { kp = 16; kprime = 16; pub_kth = 0; groups = 16; cap = 256; out_cap = 48; }
  else if (k <= 24) { kp = 32; kprime = 32; pub_kth = 0; groups = 32; cap = 256; out_cap = 80; }
  else {
    kp = 16; pub_kth = 1;
    groups = (k + k / 4 + 15) / 16;
    if (groups > 16) return fail(COLTT_ERR_UNSUPPORTED, "FAST: top-k above 200 is served by the exact path");
    kprime = groups * 16;
    cap = 512;       // two tiles of rows arrive before the first cross-column bound
    out_cap = 64;    // survivors per column at the end: K' * (a small factor) / n_cols, far below this
  }
  static const char* pair_env = getenv("COLTT_FAST_PAIR");
  const bool pair = nq > 128 && !(pair_env && atoi(pair_env) == 0);   // CTA pairs (cta_group::2) once there are two query tiles
  // stage width of the CTA-pair kernel: COLTT_FAST_SB=64|128 (experiment knob; the default is the measured better one)
  static const uint32_t env_sb = [] { const char* e = getenv("COLTT_FAST_SB"); const int v = e ? atoi(e) : 64; return v == 128 ? 128u : 64u; }();
  const uint32_t sb = pair ? env_sb : (uint32_t)kBKB;
  const size_t a_bytes = (size_t)kblocks * 128 * kBK;
  const size_t stage = (size_t)(pair ? kBN / 2 : kBN) * sb;
  const size_t misc = 2 * kBN * 4 + 256;
  const size_t total = 227 * 1024;
  if (a_bytes + misc + 2 * stage > total)
    return fail(COLTT_ERR_UNSUPPORTED, "FAST: query tile does not fit shared memory (rows wider than 1536 bytes)");
  uint32_t ns = (uint32_t)((total - a_bytes - misc) / stage);
  if (ns > 8) ns = 8;
  plan->kblocks = kblocks;
  plan->kprime = kprime;
  plan->kp = kp;
  plan->pub_kth = pub_kth;
  plan->groups = groups;
  plan->cand_cap = cap;
  plan->cand_out_cap = out_cap;
  plan->n_stages = ns;
  plan->pair = pair ? 1 : 0;
  plan->fp8 = fp8 ? 1 : 0;
  plan->sb = sb;
  if (pair) {
    plan->grid_y = (nq + 255) / 256;
    uint32_t pairs = (uint32_t)(n_sms / 2) / plan->grid_y;
    if (pairs < 1) pairs = 1;
    plan->grid_x = 2 * pairs;              // CTAs along x: (pair, rank)
    plan->n_cols = pairs;                  // CTAs that see one query (x2 epilogue sets, see gemm_filter_cols)
  } else {
    plan->grid_y = (nq + 127) / 128;
    plan->grid_x = n_sms / plan->grid_y;
    if (plan->grid_x < 1) plan->grid_x = 1;
    plan->n_cols = plan->grid_x;
  }
  plan->smem_bytes = a_bytes + (size_t)ns * stage + misc;
  plan->q_stride = kblocks * kBK;
  plan->tile_rows = kBN;
  return COLTT_OK;
}

// Tensor maps describe the operands as rows of `ew`-byte words (ew = 1, 2 or 4): swizzle patterns and box shapes are byte
// geometry, identical for fp16 and E4M3, and every extent here is a multiple of 4 bytes (stored rows are zero padded to 16
// bytes, query rows to 128), so the word width is free to choose — it only changes how many elements the TMA unit counts
// per box.  Kernel-side coordinates are in BYTES and are divided by ew on the host side of the map (p.tma_shift).
static int encode_map(CUtensorMap* tm, const void* base, uint64_t inner_bytes, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_inner_bytes,
                      uint32_t box_rows, CUtensorMapSwizzle swz, uint32_t ew) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(COLTT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {inner_bytes / ew, rows};
  const cuuint64_t gstride[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner_bytes / ew, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = ew == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : (ew == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLTT_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return COLTT_OK;
}

// number of columns (epilogue sets of the CTAs that see one query) for this plan and shard size: the [query][column]
// stride of cand_out / cand_cnt / pub
uint32_t gemm_filter_cols(const GemmPlan& plan, uint32_t n_rows) {
  const uint32_t n_tiles = (n_rows + kBN - 1) / kBN;
  return (plan.n_cols < n_tiles ? plan.n_cols : n_tiles) * kEpiSets;
}

int launch_gemm_filter(const GemmParams& p_in, const GemmPlan& plan_in, const void* d_rows, uint32_t row_stride, cudaStream_t stream,
                       GemmMapCache* cache) {
  static_assert(sizeof(CUtensorMap) == 128, "GemmMapCache holds CUtensorMap as 128 opaque bytes");
  GemmMapCache local;
  GemmMapCache& mc = cache ? *cache : local;
  CUtensorMap& tm = *reinterpret_cast<CUtensorMap*>(mc.maps[0]);
  CUtensorMap& tmq = *reinterpret_cast<CUtensorMap*>(mc.maps[1]);
  CUtensorMap& tmpf = *reinterpret_cast<CUtensorMap*>(mc.maps[2]);
  GemmPlan plan = plan_in;
  const uint32_t cols = gemm_filter_cols(plan, p_in.n_rows) / kEpiSets;   // CTAs (or pairs) along x
  if (cols * kEpiSets < plan.groups) return fail(COLTT_ERR_UNSUPPORTED, "FAST: fewer columns than bound classes");
  const uint32_t row_bytes = (p_in.dim * (plan.fp8 ? 1u : 2u) + 3) / 4 * 4;   // stored rows are zero padded to 16 bytes
  static const uint32_t env_ew = [] { const char* e = getenv("COLTT_TMA_WORD"); const int v = e ? atoi(e) : 4; return v == 1 ? 1u : (v == 2 ? 2u : 4u); }();
  const uint32_t ew = env_ew, tma_shift = ew == 4 ? 2u : (ew == 2 ? 1u : 0u);
  const uint32_t pf_inner = 256;                                           // bytes per L2-prefetch request row
  const uint32_t box_rows = plan.pair ? kBN / 2 : kBN;
  int rc;
  if (mc.rows != d_rows || mc.n_rows != p_in.n_rows || mc.row_bytes != row_bytes || mc.row_stride != row_stride || mc.box_rows != box_rows ||
      mc.pf_inner != pf_inner || mc.sb != plan.sb || mc.ew != ew) {
    mc.rows = nullptr;
    // shard tile stages: 256 (or 128 per CTA of a pair) rows x 64 B, 64B swizzle
    rc = encode_map(&tm, d_rows, row_bytes, p_in.n_rows, row_stride, plan.sb, box_rows, plan.sb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, ew);
    if (rc) return rc;
    // L2 prefetch view of the shard: whole 128-byte lines, 128 rows per request
    rc = encode_map(&tmpf, d_rows, row_bytes, p_in.n_rows, row_stride, pf_inner, box_rows, CU_TENSOR_MAP_SWIZZLE_NONE, ew);
    if (rc) return rc;
    mc.rows = d_rows; mc.n_rows = p_in.n_rows; mc.row_bytes = row_bytes; mc.row_stride = row_stride; mc.box_rows = box_rows; mc.pf_inner = pf_inner; mc.sb = plan.sb; mc.ew = ew;
  }
  if (mc.q != p_in.q_lowered || mc.nq != p_in.nq || mc.q_stride != p_in.q_stride || mc.q_ew != ew) {
    mc.q = nullptr;
    // queries: [nq][q_stride bytes], zero padded to kblocks*128 bytes; box = 128 B x 128 rows, 128B swizzle
    rc = encode_map(&tmq, p_in.q_lowered, p_in.q_stride, p_in.nq, (uint64_t)p_in.q_stride, kBK, 128, CU_TENSOR_MAP_SWIZZLE_128B, ew);
    if (rc) return rc;
    mc.q = p_in.q_lowered; mc.nq = p_in.nq; mc.q_stride = p_in.q_stride; mc.q_ew = ew;
  }
  GemmParams p = p_in;
  p.rows = static_cast<const uint8_t*>(d_rows);
  p.row_stride = row_stride;
  p.pf_inner = pf_inner;
  p.tma_shift = tma_shift;
  p.kblocks = plan.kblocks; p.kprime = plan.kprime; p.cand_cap = plan.cand_cap; p.cand_out_cap = plan.cand_out_cap; p.n_stages = plan.n_stages;
  p.pub_kth = plan.pub_kth; p.groups = plan.groups;
  if (plan.pair) {
    plan.grid_x = 2 * cols;
    return launch_gemm_filter_pair(tm, tmq, tmpf, p, plan, stream);
  }
  dim3 grid(cols, plan.grid_y);
#define COLTT_K2(KPV, F8V)                                                                          \
  {                                                                                                 \
    auto kfn = gemm_filter_kernel<KPV, F8V>;                                                        \
    { int arc = kernel_attrs(kfn, plan.smem_bytes); if (arc) return arc; }                          \
    kfn<<<grid, kGemmThreads, plan.smem_bytes, stream>>>(tm, tmq, tmpf, p);                         \
  }
  if (plan.kp == 16) { if (plan.fp8) COLTT_K2(16, true) else COLTT_K2(16, false) }
  else { if (plan.fp8) COLTT_K2(32, true) else COLTT_K2(32, false) }
#undef COLTT_K2
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
