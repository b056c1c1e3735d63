// gemm_filter.cu — K2: batched query x shard distance as a tcgen05 GEMM with a fused
// threshold top-K' epilogue (COLTT_MATH_FAST), for fp16 rows ("bf16"/f16 stores).
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
//   M = 128 queries  -> TMEM lanes.  The query tile is the A operand and lives in TENSOR MEMORY
//                       for the whole kernel (128 lanes x dim/2 columns, written once with
//                       tcgen05.st): 128 x 768 fp16 = 192 KB would not fit beside a pipeline in
//                       shared memory, but it is 384 of the 512 TMEM columns.
//   N = 64 shard rows per tile -> B operand, K-major (= the row-major shard as stored), streamed
//                       by TMA (cp.async.bulk.tensor, 128B swizzle) through an NS-stage mbarrier ring
//                       that owns almost all of shared memory (deep enough to cover HBM latency).
//   D = 128 x 64 fp32 accumulators, double-buffered in the remaining 128 TMEM columns, so the
//                       epilogue of tile i overlaps the MMAs of tile i+1.
//   Warp 0: TMA producer.  Warp 1: tcgen05.mma issuer (one elected lane).  Warps 2-5: epilogue —
//   thread = query: tcgen05.ld the 64 scores of its query, one FFMA turns each into a
//   "larger is better" key (cosine: +-dot/||row||; L2: +-(2 dot - ||row||^2)), compare with the
//   running K'-th threshold held in a register; the rare survivors go to a per-query shared-memory
//   buffer that is compacted by the warp (rank counting) when it fills.  Thresholds are shared
//   across CTAs through one atomicMax per compaction, so the total number of survivors per
//   query is ~K' ln(N/K') over the WHOLE shard, not per CTA.
// No second pass over HBM: bytes per launch = N*dim*2 (+ N*4 norms + nq*dim*2 queries).
#include <cuda.h>

#include "kernels.cuh"
#include "store.h"

namespace coltt {

static constexpr int kGemmThreads = 192;
static constexpr int kBN = 64;          // shard rows per tile (MMA N)
static constexpr int kBK = 64;          // fp16 elements per K block = one 128-byte swizzle row
static constexpr int kCandStride = 129; // padded query stride of the candidate buffers (bank-conflict free)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate); SASS: UTCHMMA
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
      "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// order-preserving float <-> uint32 (for atomicMax on thresholds that may be negative)
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(f);
#else
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// K-major, 128B-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 | version(1)<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int KP>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_filter_kernel(const __grid_constant__ CUtensorMap tmap, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t NS = p.n_stages, C = p.cand_cap, KB = p.kblocks;
  static_assert(KP == 16 || KP == 32, "K' is a register array");
  constexpr uint32_t STAGE_BYTES = kBN * kBK * 2;  // 8 KB

  uint8_t* b_stages = smem;
  float* cand_key = reinterpret_cast<float*>(smem + (size_t)NS * STAGE_BYTES);
  uint32_t* cand_row = reinterpret_cast<uint32_t*>(cand_key + (size_t)C * kCandStride);
  float* top_s = reinterpret_cast<float*>(cand_row + (size_t)C * kCandStride);   // [KP][kCandStride]
  float* sweep_s = top_s + (size_t)KP * kCandStride;                             // [KP][kCandStride]
  float* coef_a = sweep_s + (size_t)KP * kCandStride;                            // [2][kBN]
  float* coef_b = coef_a + 2 * kBN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(coef_b + 2 * kBN);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t n_tiles = (p.n_rows + kBN - 1) / kBN;
  const uint32_t q_tile0 = blockIdx.y * 128;

  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < NS; s++) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(smem_u32(tfull_bar + b), 1); mbar_init(smem_u32(tempty_bar + b), 4); }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + 2 * kBN;  // columns [128, 128 + KB*32): the query tile

  // ---- load the query tile into tensor memory (A operand): lane = query, 2 fp16 per column
  if (warp >= 2) {
    const uint32_t quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const uint32_t ql = quarter * 32 + lane;           // query within the tile == TMEM lane
    const uint32_t q = q_tile0 + ql;
    const uint4* src = reinterpret_cast<const uint4*>(p.q_f16 + (size_t)q * p.q_stride);
    for (uint32_t kb = 0; kb < KB; kb++) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint4 x = q < p.nq ? __ldg(src + kb * 8 + i) : make_uint4(0, 0, 0, 0);
        v[4 * i + 0] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
      }
      tmem_st32(tmem_a + ((quarter * 32) << 16) + kb * 32, v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      long long w_empty = 0, t_start = clock64();
      for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (uint32_t kb = 0; kb < KB; kb++, it++) {
          const uint32_t s = it % NS, ph = (it / NS) & 1;
          const long long c0 = clock64();
          while (!mbar_try_wait(smem_u32(empty_bar + s), ph ^ 1)) __nanosleep(64);   // do not steal issue slots from the epilogue
          w_empty += clock64() - c0;
          mbar_arrive_expect_tx(smem_u32(full_bar + s), STAGE_BYTES);
          tma_load_2d(smem_u32(b_stages + (size_t)s * STAGE_BYTES), &tmap, (int)(kb * kBK), (int)(t * kBN), smem_u32(full_bar + s));
        }
      }
      if (p.dbg_prof) {
        p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 0] = (unsigned long long)w_empty;
        p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 1] = (unsigned long long)(clock64() - t_start);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b F16 K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t it = 0, ti = 0;
      long long w_tempty = 0, w_full = 0, t_start = clock64();
      for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
        const uint32_t buf = ti & 1, bph = (ti >> 1) & 1;
        const long long c0 = clock64();
        while (!mbar_try_wait(smem_u32(tempty_bar + buf), bph ^ 1)) __nanosleep(20);   // epilogue drained this accumulator
        w_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (uint32_t kb = 0; kb < KB; kb++, it++) {
          const uint32_t s = it % NS, ph = (it / NS) & 1;
          const long long c1 = clock64();
          while (!mbar_try_wait(smem_u32(full_bar + s), ph)) __nanosleep(20);
          w_full += clock64() - c1;
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_stages + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (uint32_t j = 0; j < kBK / 16; j++)
            umma_f16_ts(d_tmem, tmem_a + (kb * 4 + j) * 8, make_b_desc(b_addr + j * 32), idesc, (kb | j) != 0 ? 1u : 0u);
          umma_commit(smem_u32(empty_bar + s));           // frees the smem stage when these MMAs retire
        }
        umma_commit(smem_u32(tfull_bar + buf));            // accumulator complete
      }
      if (p.dbg_prof) {
        p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 2] = (unsigned long long)w_tempty;
        p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 3] = (unsigned long long)w_full;
        p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 4] = (unsigned long long)(clock64() - t_start);
      }
    }
  } else {
    // ================= epilogue: thread = query =================
    const uint32_t quarter = warp & 3;
    const uint32_t ql = quarter * 32 + lane;
    const uint32_t q = q_tile0 + ql;
    const uint32_t et = threadIdx.x - 64;               // 0..127 among the epilogue threads
    const bool q_valid = q < p.nq;
    const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
    // Per-query state.  top_s[]: this CTA's KP best keys (sorted, descending); sweep_s[]: the KP
    // largest per-CTA maxima seen in the current sweep.  Both are shared-memory columns so the rare
    // insertions are small dynamic loops, and the per-tile hot path stays branch-free and I-cache small.
    // thr = max(top_s[KP-1], G): G = KP-th largest of the best keys the CTAs publish (KP different
    // CTAs each hold a row at least that good) — the bound tracks the shard-wide KP-th best key.
    float* top_c = top_s + ql;
    float* sweep_c_ = sweep_s + ql;
    for (int i = 0; i < KP; i++) { top_c[i * kCandStride] = NEG_INF; sweep_c_[i * kCandStride] = NEG_INF; }
    float G = NEG_INF, my_best = NEG_INF;
    uint32_t next_sweep = 1;
    const bool sweeping = q_valid && gridDim.x >= (uint32_t)KP;
    float thr = q_valid ? NEG_INF : POS_INF;
    bool overflowed = false;
    uint32_t cnt = 0;
    float* my_key = cand_key + ql;                       // [slot*kCandStride]
    uint32_t* my_row = cand_row + ql;
    float* my_pub = p.pub + (size_t)blockIdx.x * p.nq + q; // pub[cta][query]: best key this CTA has seen for the query
    uint32_t ti = 0;
    long long w_tfull = 0, t_start_e = clock64();

    // insert v into a descending sorted column of KP entries (v > last entry)
    auto sorted_insert = [&](float* col, float v) {
      int i = KP - 1;
      while (i > 0) {
        const float up = col[(i - 1) * kCandStride];
        if (up >= v) break;
        col[i * kCandStride] = up;
        --i;
      }
      col[i * kCandStride] = v;
    };

    // per-row key coefficients: key = acc * a + b, larger is better for the select mode.  The
    // ||row||^2 of tile i+1 is fetched while tile i is being processed (its latency is off the path).
    auto coef_store = [&](uint32_t b_, uint32_t row, float n2) {
      float a = 0.0f, b = NEG_INF;
      if (row < p.n_rows) {
        if (p.metric == COLTT_COSINE) {
          if (n2 > 0.0f) { a = p.nearest ? rsqrtf(n2) : -rsqrtf(n2); b = 0.0f; }
          else b = p.nearest ? NEG_INF : POS_INF;       // zero row: NaN distance, last in T order
        } else {
          a = p.nearest ? 2.0f : -2.0f;
          b = p.nearest ? -n2 : n2;
        }
      }
      coef_a[b_ * kBN + et] = a;
      coef_b[b_ * kBN + et] = b;
    };
    if (et < (uint32_t)kBN && blockIdx.x < n_tiles) {
      const uint32_t row = blockIdx.x * kBN + et;
      coef_store(0, row, row < p.n_rows ? p.row_norm2[row] : 0.0f);
    }
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ti++) {
      const uint32_t buf = ti & 1, bph = (ti >> 1) & 1;
      const uint32_t row0 = t * kBN;
      named_bar_sync(1, 128);
      const uint32_t next_row = (t + gridDim.x) * kBN + et;
      float n2_next = 0.0f;
      if (et < (uint32_t)kBN && next_row < p.n_rows) n2_next = p.row_norm2[next_row];
      const long long ce0 = clock64();
      mbar_wait(smem_u32(tfull_bar + buf), bph);
      w_tfull += clock64() - ce0;
      tc_fence_after();
#pragma unroll 1
      for (uint32_t half = 0; half < kBN / 32; half++) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((quarter * 32) << 16) + buf * kBN + half * 32, v);
        tmem_wait_ld();
        if (half == kBN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tempty_bar + buf));   // accumulator free for tile ti+2
        }
        if (p.dbg_acc && q_valid) {
#pragma unroll
          for (int c = 0; c < 32; c++) {
            const uint32_t row = row0 + half * 32 + c;
            if (row < p.n_rows) p.dbg_acc[(size_t)q * p.n_rows + row] = __uint_as_float(v[c]);
          }
        }
        if (p.dbg_flags & 1u) continue;
        // hot path: 32 FFMA + a max tree + one vote; no per-element branches
        const float4* ca = reinterpret_cast<const float4*>(coef_a + buf * kBN + half * 32);
        const float4* cb = reinterpret_cast<const float4*>(coef_b + buf * kBN + half * 32);
        float key[32];
        float kmax = NEG_INF;
#pragma unroll
        for (int c4 = 0; c4 < 8; c4++) {
          const float4 a4 = ca[c4], b4 = cb[c4];
          key[4 * c4 + 0] = fmaf(__uint_as_float(v[4 * c4 + 0]), a4.x, b4.x);
          key[4 * c4 + 1] = fmaf(__uint_as_float(v[4 * c4 + 1]), a4.y, b4.y);
          key[4 * c4 + 2] = fmaf(__uint_as_float(v[4 * c4 + 2]), a4.z, b4.z);
          key[4 * c4 + 3] = fmaf(__uint_as_float(v[4 * c4 + 3]), a4.w, b4.w);
          kmax = fmaxf(kmax, fmaxf(fmaxf(key[4 * c4 + 0], key[4 * c4 + 1]), fmaxf(key[4 * c4 + 2], key[4 * c4 + 3])));
        }
        const bool mine = kmax > thr;
        if (__any_sync(0xffffffffu, mine)) {
          // rare path: park the 32 keys in the free tail of the candidate buffer (room for 32 is an
          // invariant), then walk them with a dynamic loop
          if (mine) {
#pragma unroll
            for (int c = 0; c < 32; c++) my_key[(cnt + c) * kCandStride] = key[c];
            const uint32_t base = cnt, rb = row0 + half * 32;
            uint32_t mask = 0;
#pragma unroll
            for (int c = 0; c < 32; c++) mask |= (key[c] > thr ? 1u : 0u) << c;
            uint32_t w = cnt;
            while (mask) {
              const uint32_t c = __ffs(mask) - 1;
              mask &= mask - 1;
              const float kk = my_key[(base + c) * kCandStride];
              if (kk > thr) {   // thr may have risen since the mask was taken
                my_key[w * kCandStride] = kk;
                my_row[w * kCandStride] = rb + c;
                w++;
                if (kk > top_c[(KP - 1) * kCandStride]) {
                  sorted_insert(top_c, kk);
                  thr = fmaxf(thr, top_c[(KP - 1) * kCandStride]);
                }
                if (kk > my_best) { my_best = kk; __stcg(my_pub, kk); }
              }
            }
            cnt = w;
            // the buffer must keep room for the next 32 columns: drop what fell below the bound
            if (cnt + 32 > C) {
              uint32_t w2 = 0;
              for (uint32_t s2 = 0; s2 < cnt; s2++) {
                const float k2 = my_key[s2 * kCandStride];
                const uint32_t r2 = my_row[s2 * kCandStride];
                if (k2 >= thr) { my_key[w2 * kCandStride] = k2; my_row[w2 * kCandStride] = r2; w2++; }
              }
              cnt = w2;
              if (cnt + 32 > C) {   // more than C-32 rows tie at the bound: give this query to the exact path
                overflowed = true;
                cnt = 0;
                thr = POS_INF;
              }
            }
          }
        }
      }
      if (et < (uint32_t)kBN) coef_store(buf ^ 1, next_row, n2_next);
      // Full sweep of the per-CTA maxima on an exponential schedule (tiles 1,2,3,4,6,8,12,16,...):
      // the bound moves like 1/rows-seen, so late sweeps are rare.  Loads go out in batches of 8.
      if (sweeping && !overflowed && ti == next_sweep) {
        next_sweep = ti < 4 ? ti + 1 : ti + (ti >> 1);
        for (int i = 0; i < KP; i++) sweep_c_[i * kCandStride] = NEG_INF;
        for (uint32_t c0 = 0; c0 < gridDim.x; c0 += 8) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; i++) pv[i] = c0 + i < gridDim.x ? __ldcg(p.pub + (size_t)(c0 + i) * p.nq + q) : NEG_INF;
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (pv[i] > sweep_c_[(KP - 1) * kCandStride]) sorted_insert(sweep_c_, pv[i]);
        }
        G = fmaxf(G, sweep_c_[(KP - 1) * kCandStride]);   // KP different CTAs each hold a row at least this good
        thr = fmaxf(thr, G);
      }
    }
    if (p.dbg_prof && et == 0) {
      p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 5] = (unsigned long long)w_tfull;
      p.dbg_prof[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 6] = (unsigned long long)(clock64() - t_start_e);
    }
    // ---- hand the survivors to rerank.cu: [query][cta][slot]; publish the bound they were cut at
    if (q_valid) {
      GemmCand* out = p.cand_out + ((size_t)q * gridDim.x + blockIdx.x) * C;
      uint32_t w = 0;
      for (uint32_t s2 = 0; s2 < cnt; s2++) {
        const float kk = my_key[s2 * kCandStride];
        if (kk >= thr) { out[w].key = kk; out[w].row = my_row[s2 * kCandStride]; w++; }
      }
      p.cand_cnt[(size_t)q * gridDim.x + blockIdx.x] = overflowed ? 0xffffffffu : w;
      if (!overflowed && thr > NEG_INF) atomicMax(p.g_thr + q, f2ord(thr));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
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

int plan_gemm_filter(uint32_t dim, uint32_t nq, uint32_t k, int n_sms, GemmPlan* plan) {
  const uint32_t kblocks = (dim + kBK - 1) / kBK;
  if (kblocks * 32 + 2 * kBN > 512) return fail(COLTT_ERR_UNSUPPORTED, "FAST: query tile does not fit tensor memory (dim > 768 fp16)");
  if (k > 24) return fail(COLTT_ERR_UNSUPPORTED, "FAST: top-k above 24 is served by the exact path");
  const uint32_t kprime = k <= 10 ? 16 : 32;         // K' (register-resident per query); margin for the certificate
  const uint32_t cap = kprime + 48;                  // K' + 32 columns of headroom + 16 slack
  const size_t cand_bytes = (size_t)cap * kCandStride * 8 + 2 * (size_t)kprime * kCandStride * 4;
  const size_t misc = 2 * 2 * kBN * 4 + 1024;
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - cand_bytes - misc;
  uint32_t ns = (uint32_t)(budget / (kBN * kBK * 2));
  if (ns > 24) ns = 24;
  if (ns < 4) return fail(COLTT_ERR_UNSUPPORTED, "FAST: not enough shared memory for the pipeline");
  plan->kblocks = kblocks;
  plan->kprime = kprime;
  plan->cand_cap = cap;
  plan->n_stages = ns;
  plan->grid_y = (nq + 127) / 128;
  plan->grid_x = n_sms / plan->grid_y;
  if (plan->grid_x < 1) plan->grid_x = 1;
  plan->smem_bytes = 1024 + (size_t)ns * kBN * kBK * 2 + cand_bytes + 2 * 2 * kBN * 4 + (2 * (size_t)ns + 4) * 8 + 16;
  plan->q_stride = kblocks * kBK;
  return COLTT_OK;
}

int launch_gemm_filter(const GemmParams& p_in, const GemmPlan& plan, const void* d_rows, uint32_t row_stride, cudaStream_t stream) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(COLTT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {p_in.dim, p_in.n_rows};
  const cuuint64_t gstride[1] = {row_stride};
  const cuuint32_t box[2] = {kBK, kBN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d_rows), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLTT_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  GemmParams p = p_in;
  p.kblocks = plan.kblocks; p.kprime = plan.kprime; p.cand_cap = plan.cand_cap; p.n_stages = plan.n_stages;
  uint32_t n_tiles = (p.n_rows + kBN - 1) / kBN;
  dim3 grid(plan.grid_x < n_tiles ? plan.grid_x : n_tiles, plan.grid_y);
  if (plan.kprime == 16) {
    COLTT_CUDA(cudaFuncSetAttribute(gemm_filter_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes));
    gemm_filter_kernel<16><<<grid, kGemmThreads, plan.smem_bytes, stream>>>(tm, p);
  } else {
    COLTT_CUDA(cudaFuncSetAttribute(gemm_filter_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes));
    gemm_filter_kernel<32><<<grid, kGemmThreads, plan.smem_bytes, stream>>>(tm, p);
  }
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
