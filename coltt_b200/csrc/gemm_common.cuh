// gemm_common.cuh — tcgen05 / TMA / mbarrier PTX wrappers and the fused threshold top-K' epilogue
// shared by the one-CTA (gemm_filter.cu) and CTA-pair (gemm_filter2.cu) filter kernels.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace coltt {

static constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two sets of four)
static constexpr int kEpiThreads = 256;
static constexpr int kEpiSets = 2;         // each set owns half of every tile's columns and its own per-query state
static constexpr int kBN = 256;            // shard rows per tile (MMA N): one 128-cycle tcgen05.mma per K step
// Operand geometry is in BYTES, so the same kernels serve fp16 rows (kind::f16, 16 elements per MMA) and E4M3 rows
// (kind::f8f6f4, 32 elements per MMA): one tcgen05.mma consumes 32 bytes of K from each operand row.
static constexpr int kBK = 128;            // bytes per K block of the resident query tile (128-byte swizzle rows)
static constexpr int kBKB = 64;            // bytes per row of a shard-tile stage (64-byte swizzle rows) = 2 MMAs

// Role timers (clock64 around every wait) are compiled in only with -DCOLTT_K2_PROF=1: the reads sit on the
// single-thread MMA issue path, where they cost more than the work they measure.
#ifndef COLTT_K2_PROF
#define COLTT_K2_PROF 0
#endif
#if COLTT_K2_PROF
#define K2_NOW() clock64()
#else
#define K2_NOW() 0ll
#endif

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// CTA-pair commit: arrives on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// SS-mode: D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate); SASS: UTCHMMA
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 with both operands E4M3 (instruction-descriptor formats 0/0, the same bits as F16/F16 under kind::f16);
// SASS: UTCQMMA
__device__ __forceinline__ void umma_f8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool FP8>
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (FP8) umma_f8_ss(d, a, b, idesc, acc); else umma_f16_ss(d, a, b, idesc, acc);
}
template <bool FP8>
__device__ __forceinline__ void umma_ss_pair(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (FP8) umma_f8_ss_pair(d, a, b, idesc, acc); else umma_f16_ss_pair(d, a, b, idesc, acc);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
// CTA-pair load: data lands in THIS CTA's shared memory, completion is signalled on the barrier at `bar`
// in the pair's leader (even) CTA — the peer bit of the barrier address is cleared as CUTLASS does.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
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
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// one lane of a converged warp (warp-uniform control flow keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// K-major, 128B-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 | version(1)<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, 64B-swizzled operand (8-row atoms of 512 B): SBO = 512>>4, SWIZZLE_64B = 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

// key[c] for a per-lane dynamic c out of 32 registers: a 5-level select tree (31 SEL), no memory
__device__ __forceinline__ float pick32(const float (&k)[32], uint32_t c) {
  float a[16], b[8], d[4], e[2];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = (c & 16u) ? k[i + 16] : k[i];
#pragma unroll
  for (int i = 0; i < 8; i++) b[i] = (c & 8u) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; i++) d[i] = (c & 4u) ? b[i + 4] : b[i];
#pragma unroll
  for (int i = 0; i < 2; i++) e[i] = (c & 2u) ? d[i + 2] : d[i];
  return (c & 1u) ? e[1] : e[0];
}

// Cross-column bound G of filter_epilogue (see its header comment): reads published values of query q.  ANY subset of the
// columns gives a valid bound (fewer vouching columns only make it looser), so a sweep reads at most kSweepCols of them —
// a window that rotates from sweep to sweep — instead of all n_cols: the epilogue warps do not drain accumulators while
// they sweep, and the role timers showed the MMA issuer waiting on exactly that (7 sweeps x ~10 K cycles of 148 dependent-
// latency loads = the whole "waiting for the epilogue" share, 13 % of the kernel).  The loads of one pass (one per class) are
// issued together BEFORE any of them is consumed: each is an L2 round trip, and a sweep that serialises them (load, max,
// load, max ...) costs ~100 K cycles — measured: it doubled the kernel time.  Out of line: cold code next to the hot loop.
static constexpr uint32_t kSweepCols = 32;
template <int KP>
__device__ __noinline__ float cross_column_bound(const float* pub, uint32_t nq, uint32_t q, uint32_t col, uint32_t n_cols, uint32_t groups,
                                                 uint32_t first) {
  const float NEG_INF = __int_as_float(0xff800000);
  float gmax[KP];
#pragma unroll
  for (int i = 0; i < KP; i++) gmax[i] = NEG_INF;
  // window of columns [first, first + span) modulo n_cols; span is a multiple of `groups` so every class gets the same share
  uint32_t span = n_cols < kSweepCols ? n_cols : kSweepCols;
  if (span > groups) span -= span % groups;
#pragma unroll 2
  for (uint32_t c0 = 0; c0 < span; c0 += groups) {
    float pv[KP];
#pragma unroll
    for (int i = 0; i < KP; i++) {
      uint32_t c = first + c0 + i;
      if (c >= n_cols) c -= n_cols;
      const bool ok = (uint32_t)i < groups && c0 + i < span;
      pv[i] = __ldcg(pub + (size_t)(ok ? c : col) * nq + q);   // own column when masked: a valid address, value unused
      if (!ok) pv[i] = NEG_INF;
    }
#pragma unroll
    for (int i = 0; i < KP; i++) gmax[i] = fmaxf(gmax[i], pv[i]);   // fmaxf ignores the NaN "nothing published yet" marker
  }
  float G = gmax[0];
#pragma unroll
  for (int i = 1; i < KP; i++) if ((uint32_t)i < groups) G = fminf(G, gmax[i]);
  return G;
}

// -------------------------------------------------------------------------------------------------
// Epilogue of the filter GEMM (warps 2-9, thread = query).  For every 256-row tile: tcgen05.ld the
// 128x256 fp32 accumulator 32 columns at a time, one FFMA turns each score into a "larger is better" key
// (cosine: +-dot * s_row/||row||; L2: +-(2 dot - ||row||^2)), a max tree and ONE warp vote per 32 scores decide
// whether anything beats the running bound held in a register; the rare survivors are appended to an
// L2-resident per-(column,query) buffer.  A "column" is one epilogue set of one CTA: the columns partition
// the shard's rows.
//
// The bound `thr` must never exceed the K'-th best key of the whole shard (K' = p.kprime >= K): a row is
// dropped only when key <= thr.  Two sound sources, generalised from one scheme — every column keeps its own
// KP best keys in registers (top[], sorted) and PUBLISHES one order statistic of them; with the columns
// split into `groups` classes (column c is in class c % groups),
//     G = min over classes of (max over the class's columns of the published value)
// has at least `groups` distinct columns holding a published value >= G:
//   * p.pub_kth == 0 (K' = KP: top-10 / top-24): a column publishes its BEST key, groups = KP  => KP rows >= G;
//     its own KP-th best, top[KP-1], is a second bound (KP local rows >= it);
//   * p.pub_kth == 1 (K' = groups * KP: top-100): a column publishes its KP-th best key, so each of the `groups`
//     columns vouches for KP rows >= G  => groups * KP = K' rows >= G.  top[KP-1] alone vouches for only KP < K'
//     rows, so here thr comes from G only.
// (A wrong bound could never return a wrong answer — rerank.cu certifies the result against the bound the
// survivors were cut at and sends uncertified queries to the exact kernel — but it would cost that re-run.)
//   col / n_cols: this CTA's index among the CTAs that see this query, and how many there are
//   tile0 / tile_stride: the tiles this CTA processes
//   arrive_tempty(buf): releases accumulator `buf` to the MMA issuer (local or leader-CTA barrier)
template <int KP, class ArriveFn>
__device__ __forceinline__ void filter_epilogue(const GemmParams& p, uint32_t tmem_base, float* coef_a, float* coef_b, uint64_t* tfull_bar,
                                                ArriveFn arrive_tempty, uint32_t tile0, uint32_t tile_stride, uint32_t n_tiles,
                                                uint32_t q_tile0, uint32_t col, uint32_t n_cols, uint32_t buf_slot, uint32_t prof_slot) {
  // col / n_cols / buf_slot arrive per CTA and are refined per epilogue set below
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t C = p.cand_cap;
  const uint32_t quarter = warp & 3;                   // TMEM lane quarter this warp may touch
  const uint32_t set = (warp - 2) >> 2;                // warps 2-5: columns 0-127 of every tile; warps 6-9: 128-255
  const uint32_t ql = quarter * 32 + lane;             // query within the tile == TMEM lane
  const uint32_t q = q_tile0 + ql;
  const uint32_t et = threadIdx.x - 64;                // 0..255 among the epilogue threads
  constexpr uint32_t CHUNKS = kBN / 32 / kEpiSets;     // 32-column chunks per set per tile
  col = col * kEpiSets + set;                          // every set is its own "column" of candidates / published values
  n_cols *= kEpiSets;
  buf_slot = buf_slot * kEpiSets + set;
  const bool q_valid = q < p.nq;
  const bool pub_kth = p.pub_kth != 0;
  const uint32_t groups = p.groups;
  const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
  float top[KP];
#pragma unroll
  for (int i = 0; i < KP; i++) top[i] = NEG_INF;
  float my_best = NEG_INF;
  uint32_t next_sweep = 1;
  const bool sweeping = q_valid && n_cols >= groups;
  float thr = q_valid ? NEG_INF : POS_INF;
  bool overflowed = false;
  uint32_t cnt = 0;
  GemmCand* my_buf = p.cand_buf + ((size_t)buf_slot * C) * 128 + ql;   // [cta][slot][128 queries], written rarely
  float* my_pub = p.pub + (size_t)col * p.nq + q;                        // pub[column][query]
  uint32_t ti = 0;
  long long w_tfull = 0, t_start_e = K2_NOW(), c_bar = 0, c_hot = 0, c_slow = 0, c_sweep = 0, n_slow = 0;

  auto coef_store = [&](uint32_t idx, uint32_t row, float n2, float sc) {
    float a = 0.0f, b = NEG_INF;
    if (row < p.n_rows) {
      if (p.metric == COLTT_COSINE) {
        if (n2 > 0.0f) { a = (p.nearest ? rsqrtf(n2) : -rsqrtf(n2)) * sc; b = 0.0f; }
        else b = p.nearest ? NEG_INF : POS_INF;       // zero row: NaN distance, last in T order
      } else {
        a = p.nearest ? 2.0f : -2.0f;
        b = p.nearest ? -n2 : n2;
      }
    }
    coef_a[idx] = a;
    coef_b[idx] = b;
  };
  uint32_t sweep_first = col % n_cols;     // rotating window start: different columns vouch in different sweeps
  auto sweep = [&]() {
    thr = fmaxf(thr, cross_column_bound<KP>(p.pub, p.nq, q, col, n_cols, groups, sweep_first));
    sweep_first += kSweepCols;
    if (sweep_first >= n_cols) sweep_first %= n_cols;
  };
  // ||row||^2 (and the E4M3 row scale) of the next tile are fetched while the current one is processed (one row per epilogue thread)
  float n2_a = 0.0f, sc_a = 1.0f;
  {
    const uint32_t ra = tile0 * kBN + et;
    if (tile0 < n_tiles && ra < p.n_rows) { n2_a = p.row_norm2[ra]; if (p.row_scale) sc_a = p.row_scale[ra]; }
  }
  for (uint32_t t = tile0; t < n_tiles; t += tile_stride, ti++) {
    const uint32_t buf = ti & 1, bph = (ti >> 1) & 1;
    const uint32_t row0 = t * kBN;
    const long long cb0 = K2_NOW();
    coef_store(et, row0 + et, n2_a, sc_a);
    named_bar_sync(1, kEpiThreads);                      // coefficients of this tile visible
    c_bar += K2_NOW() - cb0;
    {
      const uint32_t ra = (t + tile_stride) * kBN + et;
      n2_a = ra < p.n_rows ? p.row_norm2[ra] : 0.0f;
      sc_a = (p.row_scale && ra < p.n_rows) ? p.row_scale[ra] : 1.0f;
    }
    const long long ce0 = K2_NOW();
    mbar_wait(smem_u32(tfull_bar + buf), bph);
    w_tfull += K2_NOW() - ce0;
    tc_fence_after();
    const uint32_t tbase = tmem_base + ((quarter * 32) << 16) + buf * kBN + set * CHUNKS * 32;
#if COLTT_K2_PROF
    const bool drain_only = (p.dbg_flags & 1u) != 0;     // pipeline-speed probe
#endif
    uint32_t v[32];
    tmem_ld32(tbase, v);
#pragma unroll 1
    for (uint32_t hc = 0; hc < CHUNKS; hc++) {
      const uint32_t half = set * CHUNKS + hc;
      const long long cl0 = K2_NOW();
      tmem_wait_ld();
      if (p.dbg_acc && q_valid) {                        // test hook: raw accumulators
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const uint32_t row = row0 + half * 32 + c;
          if (row < p.n_rows) p.dbg_acc[(size_t)q * p.n_rows + row] = __uint_as_float(v[c]);
        }
      }
      // hot path: 32 FFMA + a max tree + one vote; no per-element branches
      const float4* ca = reinterpret_cast<const float4*>(coef_a + half * 32);
      const float4* cb = reinterpret_cast<const float4*>(coef_b + half * 32);
      float key[32];
#pragma unroll
      for (int c4 = 0; c4 < 8; c4++) {
        const float4 a4 = ca[c4], b4 = cb[c4];
        key[4 * c4 + 0] = fmaf(__uint_as_float(v[4 * c4 + 0]), a4.x, b4.x);
        key[4 * c4 + 1] = fmaf(__uint_as_float(v[4 * c4 + 1]), a4.y, b4.y);
        key[4 * c4 + 2] = fmaf(__uint_as_float(v[4 * c4 + 2]), a4.z, b4.z);
        key[4 * c4 + 3] = fmaf(__uint_as_float(v[4 * c4 + 3]), a4.w, b4.w);
      }
      // v[] is dead: the next 32 columns stream out of tensor memory while this chunk is ranked
      if (hc + 1 < CHUNKS) {
        tmem_ld32(tbase + (hc + 1) * 32, v);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty(buf);               // this warp is done with the accumulator of tile ti
      }
#if COLTT_K2_PROF
      if (drain_only) continue;
#endif
      float kmax = NEG_INF;
#pragma unroll
      for (int c4 = 0; c4 < 8; c4++)
        kmax = fmaxf(kmax, fmaxf(fmaxf(key[4 * c4 + 0], key[4 * c4 + 1]), fmaxf(key[4 * c4 + 2], key[4 * c4 + 3])));
      const bool mine = kmax > thr;
      const long long ch1 = K2_NOW();
      c_hot += ch1 - cl0;
      if (__any_sync(0xffffffffu, mine)) {
        // rare path, kept small: pass mask, then each lane walks its own set bits; the key of column c
        // comes out of the 32 key registers through a select tree (no memory, no re-read)
        n_slow++;
        if (mine) {
          uint32_t mask = 0;
#pragma unroll
          for (int c = 0; c < 32; c++) mask |= (key[c] > thr ? 1u : 0u) << c;
          const uint32_t rb = row0 + half * 32;
          while (mask) {
            const uint32_t c = __ffs(mask) - 1;
            mask &= mask - 1;
            const float kk = pick32(key, c);
            if (kk > thr) {   // thr may have risen since the mask was taken
              GemmCand e; e.key = kk; e.row = rb + c;
              my_buf[(size_t)cnt * 128] = e;
              cnt++;
              if (kk > top[KP - 1]) {
                top[KP - 1] = kk;
#pragma unroll
                for (int i = KP - 1; i > 0; --i) {
                  const float hi = fmaxf(top[i - 1], top[i]), lo = fminf(top[i - 1], top[i]);
                  top[i - 1] = hi;
                  top[i] = lo;
                }
                if (!pub_kth) thr = fmaxf(thr, top[KP - 1]);                  // KP = K' local rows are at least this good
                else if (top[KP - 1] > NEG_INF) __stcg(my_pub, top[KP - 1]);  // this column vouches for KP rows >= it
              }
              if (!pub_kth && kk > my_best) { my_best = kk; __stcg(my_pub, kk); }
            }
          }
          // the buffer must keep room for the next 32 columns: refresh the cross-column bound, then drop what fell
          // below it (rare: the buffer is sized so that a typical scan never fills it); loads go out 8 at a time
          if (cnt + 32 > C) {
            if (sweeping) sweep();
            uint32_t w2 = 0;
            for (uint32_t s0 = 0; s0 < cnt; s0 += 8) {
              GemmCand e8[8];
#pragma unroll
              for (int u = 0; u < 8; u++) if (s0 + u < cnt) e8[u] = my_buf[(size_t)(s0 + u) * 128];   // written by this thread only
#pragma unroll
              for (int u = 0; u < 8; u++)
                if (s0 + u < cnt && e8[u].key >= thr) { my_buf[(size_t)w2 * 128] = e8[u]; w2++; }
            }
            cnt = w2;
            if (cnt + 32 > C) {   // more than C-32 rows at or above the bound: give this query to the exact path
              overflowed = true;
              cnt = 0;
              thr = POS_INF;
            }
          }
        }
        c_slow += K2_NOW() - ch1;
      }
    }
    const long long cs0 = K2_NOW();
    named_bar_sync(2, kEpiThreads);                      // everyone is done reading this tile's coefficients
    c_bar += K2_NOW() - cs0;
    // Cross-column bound on a doubling schedule (tiles 1,2,4,8,...): the bound moves like 1/rows-seen.
    if (sweeping && !overflowed && ti == next_sweep) {
      const long long cw0 = K2_NOW();
      next_sweep = ti * 2;
      sweep();
      c_sweep += K2_NOW() - cw0;
    }
  }
#if COLTT_K2_PROF
  if (p.dbg_prof && et == 0) {   // set 0, quarter 2's first lane
    p.dbg_prof[(size_t)prof_slot * 8 + 5] = (unsigned long long)w_tfull;
    p.dbg_prof[(size_t)prof_slot * 8 + 6] = (unsigned long long)(K2_NOW() - t_start_e);
    p.dbg_prof[(size_t)prof_slot * 8 + 7] = (unsigned long long)c_bar;
    p.dbg_prof2[(size_t)prof_slot * 8 + 0] = (unsigned long long)0;
    p.dbg_prof2[(size_t)prof_slot * 8 + 1] = (unsigned long long)c_hot;
    p.dbg_prof2[(size_t)prof_slot * 8 + 2] = (unsigned long long)c_sweep;
    p.dbg_prof2[(size_t)prof_slot * 8 + 3] = (unsigned long long)n_slow;
    p.dbg_prof2[(size_t)prof_slot * 8 + 4] = (unsigned long long)cnt;
    p.dbg_prof2[(size_t)prof_slot * 8 + 5] = (unsigned long long)c_slow;
  }
#else
  (void)w_tfull; (void)t_start_e; (void)c_bar; (void)c_hot; (void)c_slow; (void)c_sweep; (void)n_slow; (void)prof_slot;
#endif
  // ---- hand the survivors to rerank.cu: [query][column][slot]; publish the bound they were cut at
  if (sweeping && !overflowed) sweep();   // the other columns have published more since the last scheduled sweep
  if (q_valid) {
    const uint32_t CO = p.cand_out_cap;
    GemmCand* out = p.cand_out + ((size_t)q * n_cols + col) * CO;
    uint32_t w = 0;
    for (uint32_t s0 = 0; s0 < cnt; s0 += 8) {
      GemmCand e8[8];
#pragma unroll
      for (int u = 0; u < 8; u++) if (s0 + u < cnt) e8[u] = my_buf[(size_t)(s0 + u) * 128];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (s0 + u < cnt && e8[u].key >= thr) { if (w < CO) out[w] = e8[u]; w++; }
    }
    if (w > CO) overflowed = true;   // more survivors than the hand-off slot holds: exact path for this query
    p.cand_cnt[(size_t)q * n_cols + col] = overflowed ? 0xffffffffu : w;
    if (!overflowed && thr > NEG_INF) atomicMax(p.g_thr + q, f2ord(thr));
  }
}

}  // namespace coltt
