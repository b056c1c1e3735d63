// gemm_filter2.cu — K2, CTA-pair variant (tcgen05 cta_group::2) for batches of more than 128 queries.
//
// Same role as gemm_filter.cu (the batched query x shard filter behind COLTT_MATH_FAST), but two CTAs of
// one thread-block cluster work as a pair on every shard tile:
//   * M = 256 queries: CTA rank r keeps queries [256*y + 128*r, +128) resident in its shared memory and
//     gets the matching 128 x 256 accumulator in its own tensor memory;
//   * every 256-row shard tile is loaded ONCE for both query halves: each CTA TMA-loads 128 of the 256 rows
//     (8 KB per 32-element K stage) and tcgen05.mma.cta_group::2 reads both halves — per SM the L2->SM
//     traffic and the shared-memory operand reads of B are halved (ncu on the one-CTA kernel showed an
//     M128xN256 SS MMA costs ~280 cycles against the 128-cycle math floor, shared-memory-read bound);
//   * the leader (rank 0) issues the MMAs; tcgen05.commit multicasts "stage free" / "accumulator ready"
//     to both CTAs' mbarriers; both CTAs' epilogue warps release accumulators on the leader's barrier.
// Warp roles per CTA as in gemm_filter.cu; the epilogue is the shared filter_epilogue().
#include "gemm_common.cuh"
#include "store.h"

namespace coltt {

// SB = bytes per row of a shard-tile stage: 64 (64B swizzle, 2 MMAs per stage) or 128 (128B swizzle, 4 MMAs per stage
// and half as many barrier round trips per byte; chosen by the plan)
template <int KP, bool FP8, int SB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_filter_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ CUtensorMap tmap_pf, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t NS = p.n_stages, KB = p.kblocks;
  constexpr uint32_t HALF_ROWS = kBN / 2;                   // shard rows each CTA of the pair loads
  constexpr uint32_t STAGE_BYTES = HALF_ROWS * SB;          // 128 rows x 64 B = 8 KB (or x 128 B = 16 KB) per CTA per stage
  constexpr uint32_t ABLK_BYTES = 128 * kBK;
  constexpr uint32_t MPS = SB / 32;                         // tcgen05.mma per stage
  const uint32_t NSTEP = KB * (kBK / SB);

  uint8_t* a_smem = smem;
  uint8_t* b_stages = a_smem + (size_t)KB * ABLK_BYTES;     // [NS][128 rows][64 B], 64B-swizzled
  float* coef_a = reinterpret_cast<float*>(b_stages + (size_t)NS * STAGE_BYTES);
  float* coef_b = coef_a + kBN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(coef_b + kBN);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tfull_bar = empty_bar + NS;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;   // [2]   (used in the leader only)
  uint64_t* a_bar = tempty_bar + 2;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t rank = cluster_ctarank();                  // 0 = leader
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t n_tiles = (p.n_rows + kBN - 1) / kBN;
  const uint32_t q_tile0 = blockIdx.y * 256 + rank * 128;
  const uint32_t cta_lin = blockIdx.y * gridDim.x + blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < NS; s++) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(smem_u32(tfull_bar + b), 1); mbar_init(smem_u32(tempty_bar + b), 16); }  // 8 warps x 2 CTAs
    mbar_init(smem_u32(a_bar), 1);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pf) : "memory");
  }
  if (warp == 1) {   // the same warp of both CTAs allocates all 512 columns in both tensor memories
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // each CTA's own query half: resident A operand
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(smem_u32(a_bar), KB * ABLK_BYTES);
    for (uint32_t kb = 0; kb < KB; kb++)
      tma_load_2d(smem_u32(a_smem + (size_t)kb * ABLK_BYTES), &tmap_q, (int)((kb * kBK) >> p.tma_shift), (int)q_tile0, smem_u32(a_bar));
    mbar_wait(smem_u32(a_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // barriers initialised and both query halves resident before anything crosses CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own half of every shard tile) =================
    // The ring (4 x 8 KB per CTA) is shorter than the L2 round trip, so the time from "stage free" to "load
    // issued" is on the kernel's critical path exactly like the MMA issue loop: the whole warp walks the loop
    // (warp-uniform values stay in uniform registers), one elected lane issues, the load goes out first and
    // the L2 prefetch of the next tile (same K slice, one tile ahead) after it.
    uint32_t s = 0, ph = 0;
    const int row_off = (int)(rank * HALF_ROWS);
#if COLTT_K2_PROF
    const bool do_pf = (p.dbg_flags & 2u) == 0;
#else
    constexpr bool do_pf = true;
#endif
    const uint32_t pf_mask = p.pf_inner / SB - 1;            // pf_inner / SB is a power of two
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(b_stages);
    if (do_pf && pair < n_tiles && elect_one())
      for (uint32_t st = 0; st < NSTEP; st += pf_mask + 1) tma_prefetch_2d(&tmap_pf, (int)((st * SB) >> p.tma_shift), (int)(pair * kBN) + row_off);
    __syncwarp();
    // Batches of more than 256 queries run gridDim.y query-tile groups side by side; pair p of every group walks the SAME
    // tiles in the same order, so a tile fetched from HBM by one group is usually an L2 hit for the others: ncu measures 1.9 HBM
    // passes per 1024-query batch instead of 4.  Forcing the groups into lock-step (a window of 4 or 8 tile steps on a progress
    // counter per pair) brought that to 1.13 passes but cost 24 % of the kernel time on the power-capped, tensor-bound config 4
    // (profiles/r2_gemm_filter_pair_summary.md), so the groups run free.
#if COLTT_K2_PROF
    const uint32_t t_first = (p.dbg_flags & 4u) ? n_tiles : pair;   // probe bit 2: MMA cadence without any TMA traffic (operands = whatever is in smem)
#else
    const uint32_t t_first = pair;
#endif
    for (uint32_t t = t_first; t < n_tiles; t += n_pairs) {
      const int row = (int)(t * kBN) + row_off, row_pf = row + (int)(n_pairs * kBN);
      const bool pf = do_pf && t + n_pairs < n_tiles;
      for (uint32_t st = 0; st < NSTEP; st++) {
        mbar_wait(empty0 + s * 8, ph ^ 1);
        if (elect_one()) {
          // the leader's barrier collects both halves: it expects 2 x STAGE_BYTES, each CTA's load signals it
          if (rank == 0) mbar_arrive_expect_tx(full0 + s * 8, 2 * STAGE_BYTES);
          tma_load_2d_pair(stage0 + s * STAGE_BYTES, &tmap, (int)((st * SB) >> p.tma_shift), row, full0 + s * 8);
          if (pf && (st & pf_mask) == 0) tma_prefetch_2d(&tmap_pf, (int)((st * SB) >> p.tma_shift), row_pf);
        }
        __syncwarp();
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    // This loop is the single-thread critical path of the kernel: ring position, phase and both operand
    // descriptors advance by constant adds (no division, no descriptor rebuild, no timers) so that one
    // iteration (wait, 2 x tcgen05.mma, commit) issues in well under the 2 x 128 cycles the MMAs take.
    if (rank == 0) {
      // M = 256 (M>>4 = 16 at bit 24), N = 256
      const uint32_t idesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((256u >> 4) << 24);
      uint32_t s = 0, ph = 0, ti = 0;
      long long w_tempty = 0, w_full = 0, t_start = K2_NOW();
      const uint64_t a_desc0 = make_desc_sw128(smem_u32(a_smem));
      const uint64_t b_desc0 = SB == 128 ? make_desc_sw128(smem_u32(b_stages)) : make_desc_sw64(smem_u32(b_stages));
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      for (uint32_t t = pair; t < n_tiles; t += n_pairs, ti++) {
        const uint32_t buf = ti & 1, bph = (ti >> 1) & 1;
        const long long c0 = K2_NOW();
        mbar_wait(smem_u32(tempty_bar + buf), bph ^ 1);    // both CTAs' epilogues drained it
        w_tempty += K2_NOW() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        uint64_t a_desc = a_desc0;                           // (start address >> 4) lives in the low 14 bits
        if (SB == 64 && NS == 4 && (NSTEP & 3u) == 0) {
          // Fast issue loop (the configuration every wide-row shape runs: 4 x 8 KB stages).  The probes of round 2
          // (profiles/r2_k2_probes.md) showed that THIS loop, not the tensor pipe or the TMA stream, set the kernel's pace:
          // ~145 busy cycles per MMA on the issuing thread against the 128-cycle math floor.  So: the ring position is a
          // compile-time constant inside a 4-stage group (barrier addresses and B descriptors become immediates, one phase
          // flip per group), and the "stage full?" answer for stage u+1 is REQUESTED (mbarrier.test_wait, non-blocking)
          // before the MMAs of stage u are issued and only consumed after them, which takes the ~90-cycle probe latency
          // off the critical path; the blocking wait remains as the fallback when the answer is "not yet".
          constexpr uint64_t ABLK16 = ABLK_BYTES >> 4;
          bool ready = mbar_test_wait(full0, ph);
          for (uint32_t st = 0; st < NSTEP; st += 4) {
#pragma unroll
            for (uint32_t u = 0; u < 4; u++) {
#if COLTT_K2_PROF
              if (p.dbg_flags & 4u) ready = true;                                  // probe bit 2: no TMA traffic, nothing to wait for
#endif
              if (!ready) mbar_wait(full0 + u * 8, ph);
              // probe the next stage now (next group's first stage has the flipped phase)
              const bool last_of_tile = st + 4 >= NSTEP && u == 3;
              ready = last_of_tile ? mbar_test_wait(full0, ph ^ 1) : mbar_test_wait(full0 + ((u + 1) & 3) * 8, u == 3 ? ph ^ 1 : ph);
              tc_fence_after();
              const uint64_t ad = a_desc + (u >> 1) * ABLK16 + (u & 1) * 4;       // K blocks of 128 B, two 64-byte stages each
              const uint64_t bd = b_desc0 + (uint64_t)(u * (STAGE_BYTES >> 4));
              if (elect_one()) {
#if COLTT_K2_PROF
                if (!(p.dbg_flags & 8u)) {                                         // probe bit 3: the TMA stream without any MMA
#endif
                umma_ss_pair<FP8>(d_tmem, ad, bd, idesc, (st | u) != 0 ? 1u : 0u);
                umma_ss_pair<FP8>(d_tmem, ad + 2, bd + 2, idesc, 1u);             // +32 B of K
#if COLTT_K2_PROF
                }
#endif
                umma_commit_pair(empty0 + u * 8, 3);                               // stage free in both CTAs
              }
              __syncwarp();
            }
            a_desc += 2 * ABLK16;
            ph ^= 1;
          }
        } else {
        for (uint32_t st = 0; st < NSTEP; st++) {
          const long long c1 = K2_NOW();
#if COLTT_K2_PROF
          if (!(p.dbg_flags & 4u))
#endif
          mbar_wait(full0 + s * 8, ph);
          w_full += K2_NOW() - c1;
          tc_fence_after();
          const uint64_t b_desc = b_desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
          if (elect_one()) {
#if COLTT_K2_PROF
            if (!(p.dbg_flags & 8u)) {                         // probe bit 3: TMA streaming speed without any MMA
#endif
            umma_ss_pair<FP8>(d_tmem, a_desc, b_desc, idesc, st != 0 ? 1u : 0u);
#pragma unroll
            for (uint32_t m = 1; m < MPS; m++) umma_ss_pair<FP8>(d_tmem, a_desc + 2 * m, b_desc + 2 * m, idesc, 1u);   // +32 B of K each
#if COLTT_K2_PROF
            }
#endif
            umma_commit_pair(empty0 + s * 8, 3);             // stage free in both CTAs
          }
          __syncwarp();
          if (SB == 128) a_desc += (uint64_t)(ABLK_BYTES >> 4);              // one whole K block per stage
          else a_desc += (st & 1) ? (uint64_t)((ABLK_BYTES - 64) >> 4) : 4ull;    // +64 B inside a K block, then the next block
          if (++s == NS) { s = 0; ph ^= 1; }
        }
        }
        if (elect_one()) umma_commit_pair(smem_u32(tfull_bar + buf), 3);   // accumulator ready in both CTAs
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
    }
  } else {
    auto arrive = [&](uint32_t buf) { mbar_arrive_cluster(smem_u32(tempty_bar + buf), 0); };   // on the leader's barrier
    filter_epilogue<KP>(p, tmem_base, coef_a, coef_b, tfull_bar, arrive, pair, n_pairs, n_tiles, q_tile0, pair, n_pairs, cta_lin, cta_lin);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int launch_gemm_filter_pair(const CUtensorMap& tm, const CUtensorMap& tmq, const CUtensorMap& tmpf, const GemmParams& p, const GemmPlan& plan,
                            cudaStream_t stream) {
  dim3 grid(plan.grid_x, plan.grid_y);
#define COLTT_K2P(KPV, F8V, SBV)                                                                    \
  {                                                                                                 \
    auto kfn = gemm_filter_pair_kernel<KPV, F8V, SBV>;                                              \
    { int arc = kernel_attrs(kfn, plan.smem_bytes); if (arc) return arc; }                          \
    kfn<<<grid, kGemmThreads, plan.smem_bytes, stream>>>(tm, tmq, tmpf, p);                         \
  }
#define COLTT_K2P_SB(KPV, F8V) { if (plan.sb == 128) COLTT_K2P(KPV, F8V, 128) else COLTT_K2P(KPV, F8V, 64) }
  if (plan.kp == 16) { if (plan.fp8) COLTT_K2P_SB(16, true) else COLTT_K2P_SB(16, false) }
  else { if (plan.fp8) COLTT_K2P_SB(32, true) else COLTT_K2P_SB(32, false) }
#undef COLTT_K2P_SB
#undef COLTT_K2P
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
