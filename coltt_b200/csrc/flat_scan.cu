// flat_scan.cu — K1: exact-order FLAT scan with fused warp top-K (+ K3 gather front-end).
//
// Replaces the hot loop of VertexSearch / FilterableVertexSearch
// (edge/none_vectorstore.go:136-147,182-253 and the bf16/f16/f8 twins): for every stored row,
// Quantization.Similarity (edge/quantization.go:43-45, {f16,bf16,f8}_quantization.go:33-43)
// -> distance.Space.Distance (pkg/distance/space.go:61-63,93-95) -> the AVX kernel
// (pkg/distance/simd/cpp/avx.cpp:15-32,51-75) -> PriorityQueue.Add (edge/priority_queue.go:46-55).
//
// Bit-exactness: the AVX kernel keeps 8 lane accumulators (element i -> lane i%8, increasing i,
// unfused mul+add), reduces them ((l0+l1)+(l2+l3))+((l4+l5)+(l6+l7)) and then adds the scalar
// tail.  Here two threads own one row: thread g keeps lanes 4g..4g+3 as four independent
// register chains and walks the row in the same order, so every partial sum equals the
// reference's; the two halves meet through one shuffle.  Quantized rows are decoded in
// registers (fp16 -> fp32 is exact).  ||row||^2 comes precomputed from ingest (prep.cu).
//
// Data movement: HBM-bound at one query.  Each warp runs its own producer/consumer ring:
// 16 rows x `chunk_bytes` per stage, filled by 16 one-dimensional bulk async copies
// (cp.async.bulk -> UBLKCP, the TMA path: no LSU wavefronts, no register staging), completion
// on an mbarrier.  Rows land with a 16-byte pad so the 8 rows a quarter-warp reads together
// sit in distinct banks.  Row chunks stream through while the chain accumulators stay in
// registers, so the stage size is independent of dim.  With `subset` the source address of
// each copy comes from a slot list: that is the whole gather front-end (filtered search, HNSW).
//
// Algorithmic bytes per launch:  n_items*dim*elem + n_items*4 (norms)
//   + nq*dim*4 (queries) + grid*nq*k*16 (lists) — one HBM pass over the shard per QT queries.
#include <cstdlib>

#include "exact_math.cuh"
#include "kernels.cuh"
#include "store.h"
#include "topk.cuh"

namespace coltt {

static constexpr int kScanWarps = 8;
static constexpr int kRowsPerWarp = 16;
static constexpr uint32_t kRowPad = 16;
static constexpr uint32_t kF8LutFloats = 0x84 * 32;   // (code & 0x83) x 32 lane-private copies = 16.5 KB

template <int ELEM, int METRIC, int QT>
__global__ void __launch_bounds__(kScanWarps * 32, 1) flat_scan_kernel(ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float lut_tab[ELEM == ELEM_F8C ? kF8LutFloats : 1];   // F8C: per-lane copies of the decode table (exact_math.cuh)
  const float* lut_s = lut_tab + (threadIdx.x & 31);
  __shared__ int warp_cnt_s[kScanWarps][QT];
  __shared__ uint32_t cta_kth_s[QT];   // best K-th bound any warp of this CTA has reached, order-encoded

  constexpr uint32_t ES = ELEM == ELEM_F32 ? 4 : (ELEM == ELEM_F16 ? 2 : 1);   // F8C and F8E: one byte
  const uint32_t W = kScanWarps, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = p.n_stages, CB = p.chunk_bytes, RS = CB + kRowPad;
  const uint32_t stage_bytes = kRowsPerWarp * RS;

  // ---- shared memory carve-up ---------------------------------------------------------
  float* q_s = reinterpret_cast<float*>(smem);                        // [QT][q_stride]
  float* qn_s = q_s + (size_t)QT * p.q_stride;                        // [QT] (+pad to 8 floats)
  uint64_t* bars = reinterpret_cast<uint64_t*>(qn_s + 8);             // [W*S]
  // offsets only (no pointer->integer round trip): the compiler keeps the shared address space and emits LDS
  const uint32_t stages_off = (uint32_t)(((size_t)QT * p.q_stride * 4 + 32 + (size_t)W * S * 8 + 127) / 128 * 128);
  uint8_t* stages = smem + stages_off;                                // [W][S][16][RS]

  if (ELEM == ELEM_F8C)
    for (uint32_t i = threadIdx.x; i < kF8LutFloats; i += blockDim.x) lut_tab[i] = __uint_as_float(f8_compat_decode_bits((uint8_t)(i >> 5)));
  // Query groups of QT: grid.y groups run side by side, each CTA loops over the rest.  With `n_active`
  // the number of queries lives on the device (the FAST path's exact re-run of uncertified queries is
  // enqueued unconditionally and costs one empty wave when there is nothing to do); `q_map` then says
  // which prepared query each compact index stands for;
  // `q_base` windows the compact list, so scratch lists stay bounded)
  uint32_t nq_eff = p.nq;
  if (p.n_active) {
    const uint32_t na = *p.n_active;
    nq_eff = na > p.q_base ? na - p.q_base : 0u;
    if (nq_eff > p.nq) nq_eff = p.nq;
  }
  for (uint32_t q0 = blockIdx.y * QT; q0 < nq_eff; q0 += gridDim.y * QT) {
  const uint32_t nqt = nq_eff - q0 < (uint32_t)QT ? nq_eff - q0 : (uint32_t)QT;
  for (uint32_t i = threadIdx.x; i < QT * p.q_stride; i += blockDim.x) {
    uint32_t qi = i / p.q_stride, d = i - qi * p.q_stride;
    const uint32_t qsrc = qi < nqt ? (p.q_map ? p.q_map[p.q_base + q0 + qi] : q0 + qi) : 0u;
    q_s[i] = qi < nqt ? p.queries[(size_t)qsrc * p.q_stride + d] : 0.0f;
  }
  if (threadIdx.x < 8) {
    const uint32_t qsrc = threadIdx.x < nqt ? (p.q_map ? p.q_map[p.q_base + q0 + threadIdx.x] : q0 + threadIdx.x) : 0u;
    qn_s[threadIdx.x] = (threadIdx.x < nqt && METRIC == COLTT_COSINE) ? p.q_norm2[qsrc] : 0.0f;
  }
  if (lane == 0)
    for (uint32_t s = 0; s < S; s++) mbar_init(smem_u32(bars + warp * S + s), 1);
  if (threadIdx.x < QT) cta_kth_s[threadIdx.x] = p.nearest ? 0xffffffffu : 0u;   // "no bound yet"
  fence_mbar_init();
  __syncthreads();

  // ---- per-warp work list: row groups gw, gw+TW, ... each split into n_chunks stages ---
  const uint32_t row_bytes = p.row_stride;
  const uint32_t n_chunks = (row_bytes + CB - 1) / CB;
  const uint32_t CBE = CB / ES;
  const uint32_t n_groups = (p.n_items + kRowsPerWarp - 1) / kRowsPerWarp;
  const uint32_t TW = gridDim.x * W, gw = blockIdx.x * W + warp;
  const uint32_t my_groups = gw < n_groups ? (n_groups - gw + TW - 1) / TW : 0;
  const uint32_t n_work = my_groups * n_chunks;
  uint8_t* my_stages = stages + (size_t)warp * S * stage_bytes;
  const uint32_t bar0 = smem_u32(bars + warp * S);

  auto issue = [&](uint32_t t) {
    const uint32_t gi = t / n_chunks, c = t - gi * n_chunks;
    const uint32_t s = t % S;
    const uint32_t bar = bar0 + s * 8;
    const uint32_t off = c * CB;
    const uint32_t bytes = row_bytes - off < CB ? row_bytes - off : CB;
    const uint32_t item0 = (gw + gi * TW) * kRowsPerWarp;
    const uint32_t nvalid = p.n_items - item0 < (uint32_t)kRowsPerWarp ? p.n_items - item0 : (uint32_t)kRowsPerWarp;
    if (lane == 0) mbar_arrive_expect_tx(bar, bytes * nvalid);
    __syncwarp();
    if (lane < nvalid) {
      const uint32_t item = item0 + lane;
      const size_t slot = p.subset ? (size_t)p.subset[item] : (size_t)item;
      bulk_g2s(smem_u32(my_stages + (size_t)s * stage_bytes + lane * RS), p.rows + slot * p.row_stride + off, bytes, bar);
    }
  };

  // lane -> (row, half): the 8 lanes of each quarter-warp read 8 different rows (distinct banks)
  const uint32_t r = (lane & 7) + 8 * (lane >> 4);
  const uint32_t g = (lane >> 3) & 1;
  const uint32_t full8 = (p.dim / 8) * 8;

  float acc[QT][4];
  uint32_t cnt[QT];
  float kth[QT];
#pragma unroll
  for (int qi = 0; qi < QT; qi++) { cnt[qi] = 0; kth[qi] = 0.0f; }

  for (uint32_t t = 0; t < S && t < n_work; t++) issue(t);

  uint32_t gi = 0, c = 0;
  float nb_pref = 0.0f, macc_pref = 0.0f, sc_pref = 1.0f;   // sc_pref: the row's E4M3 scale
  uint32_t slot_pref = 0;
  for (uint32_t t = 0; t < n_work; t++) {
    if (c == 0) {
#pragma unroll
      for (int qi = 0; qi < QT; qi++) acc[qi][0] = acc[qi][1] = acc[qi][2] = acc[qi][3] = 0.0f;
      // ||row||^2 and the slot of this lane's row: fetched now, needed when the last chunk is done
      const uint32_t item_p = (gw + gi * TW) * kRowsPerWarp + r;
      if (item_p < p.n_items) {
        slot_pref = p.subset ? p.subset[item_p] : item_p;
        if (METRIC == COLTT_COSINE) nb_pref = p.row_norm2[slot_pref];
        if (ELEM == ELEM_F8E) sc_pref = p.row_scale[slot_pref];
        if (p.multi_acc && !p.multi_first) macc_pref = p.multi_acc[slot_pref];
      }
    }
    const uint32_t s = t % S;
    mbar_wait(bar0 + s * 8, (t / S) & 1);

    const uint8_t* rowp = my_stages + (size_t)s * stage_bytes + r * RS;
    const uint32_t e0 = c * CBE;
    const uint32_t e1 = e0 + CBE < full8 ? e0 + CBE : full8;
    const uint32_t n8 = e1 > e0 ? (e1 - e0) / 8 : 0;
    const float* qb = q_s + e0 + 4 * g;
    const uint8_t* rb = rowp + 4 * g * ES;
#pragma unroll 4
    for (uint32_t i8 = 0; i8 < n8; i8++) {
      float rv[4];
      load4<ELEM>(rb + i8 * 8 * ES, lut_s, rv);
      if (ELEM == ELEM_F8E && METRIC != COLTT_COSINE) {   // (q - s*d)^2 needs the dequantized value itself
        rv[0] = mul_rn(rv[0], sc_pref); rv[1] = mul_rn(rv[1], sc_pref); rv[2] = mul_rn(rv[2], sc_pref); rv[3] = mul_rn(rv[3], sc_pref);
      }
#pragma unroll
      for (int qi = 0; qi < QT; qi++) {
        const float4 qv = *reinterpret_cast<const float4*>(qb + (size_t)qi * p.q_stride + i8 * 8);
        if (METRIC == COLTT_COSINE) {
          acc[qi][0] = dot_step<ELEM>(acc[qi][0], qv.x, rv[0]);  // avx.cpp:60 dot += v1*v2
          acc[qi][1] = dot_step<ELEM>(acc[qi][1], qv.y, rv[1]);
          acc[qi][2] = dot_step<ELEM>(acc[qi][2], qv.z, rv[2]);
          acc[qi][3] = dot_step<ELEM>(acc[qi][3], qv.w, rv[3]);
        } else {
          float d0 = sub_rn(qv.x, rv[0]), d1 = sub_rn(qv.y, rv[1]), d2 = sub_rn(qv.z, rv[2]), d3 = sub_rn(qv.w, rv[3]);
          acc[qi][0] = add_rn(acc[qi][0], mul_rn(d0, d0));        // avx.cpp:21-23
          acc[qi][1] = add_rn(acc[qi][1], mul_rn(d1, d1));
          acc[qi][2] = add_rn(acc[qi][2], mul_rn(d2, d2));
          acc[qi][3] = add_rn(acc[qi][3], mul_rn(d3, d3));
        }
      }
    }

    if (c == n_chunks - 1) {
      // ---- row finished: reduce, tail, distance, top-K ---------------------------------
      const uint32_t item = (gw + gi * TW) * kRowsPerWarp + r;
      const bool valid = item < p.n_items;
      const uint32_t slot = valid ? slot_pref : 0u;
      const float nb = (METRIC == COLTT_COSINE && valid) ? nb_pref : 0.0f;
#pragma unroll
      for (int qi = 0; qi < QT; qi++) {
        float h = add_rn(add_rn(acc[qi][0], acc[qi][1]), add_rn(acc[qi][2], acc[qi][3]));  // avx.cpp:3-8
        float o = __shfl_xor_sync(0xffffffffu, h, 8);
        float tot = g == 0 ? add_rn(h, o) : add_rn(o, h);
        for (uint32_t d = full8; d < p.dim; d++) {  // scalar tail, avx.cpp:27-31 / :68-72
          float rv = load1<ELEM>(rowp, d - e0, lut_s);
          if (ELEM == ELEM_F8E && METRIC != COLTT_COSINE) rv = mul_rn(rv, sc_pref);
          float qv = q_s[(size_t)qi * p.q_stride + d];
          if (METRIC == COLTT_COSINE) tot = dot_step<ELEM>(tot, qv, rv);
          else { float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
        }
        // E4M3 rows: the dot product ran over the unscaled decoded values; the row's power-of-two scale multiplies every
        // product and every partial sum exactly (no over/underflow: scales are clamped to 2^+-40), so applying it once
        // here yields the bits of sum(q_i * (s*d_i)) in the reference order
        if (ELEM == ELEM_F8E && METRIC == COLTT_COSINE) tot = mul_rn(tot, sc_pref);
        float score = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn_s[qi], nb) : sqrt_via_f64(tot);
        if (p.multi_acc) {
          // score += scoreHelper(sim) * (float32(ratio) / 100), unfused, in request order (multi_vector_vertex.go:110-115)
          score = add_rn(p.multi_first ? 0.0f : macc_pref, mul_rn(score_helper(score, METRIC), p.multi_w));
          if (!p.multi_last) {
            if (valid && g == 0) p.multi_acc[slot] = score;
            continue;
          }
        }
        {   // another warp of this CTA may already hold K rows better than ours: adopt its bound
          const uint32_t shared_bits = cta_kth_s[qi];
          if (p.nearest ? shared_bits != 0xffffffffu : shared_bits != 0u) {
            const float sb = ord2f(shared_bits);
            if (cnt[qi] < p.k) { /* own list not full: keep filling it, the merge needs best-first prefixes */ }
            else kth[qi] = p.nearest ? fminf(kth[qi], sb) : fmaxf(kth[qi], sb);
          }
        }
        const bool pass = valid && g == 0 && (uint32_t)qi < nqt && (cnt[qi] < p.k || maybe_enters(score, kth[qi], p.nearest));
        uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m) {
          Hit* L = p.warp_lists + ((size_t)gw * p.nq + (q0 + qi)) * p.k;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float sc = __shfl_sync(0xffffffffu, score, src);
            const uint32_t sl = __shfl_sync(0xffffffffu, slot, src);
            const uint64_t id = p.ids[sl];
            warp_list_insert(L, p.k, cnt[qi], kth[qi], sc, sl, id, p.nearest);
          }
          if (cnt[qi] == p.k && lane == 0) {
            if (p.nearest) atomicMin(&cta_kth_s[qi], f2ord(kth[qi]));
            else atomicMax(&cta_kth_s[qi], f2ord(kth[qi]));
          }
        }
      }
    }

    __syncwarp();
    fence_proxy_async();  // our generic-proxy reads of stage s precede the async-proxy refill
    if (t + S < n_work) issue(t + S);
    if (++c == n_chunks) { c = 0; gi++; }
  }

  // ---- CTA-level merge of the W warp lists (the shard-queue merge, none_vectorstore.go:173-178)
  if (lane == 0)
#pragma unroll
    for (int qi = 0; qi < QT; qi++) warp_cnt_s[warp][qi] = (int)cnt[qi];
  __syncthreads();  // all warps done: every issued stage was consumed, stage memory is free
  // retire this group's mbarriers: the merge staging below overlays them and the next query group re-initialises the ring
  if (lane == 0)
    for (uint32_t s = 0; s < S; s++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bars + warp * S + s)) : "memory");
  __syncthreads();
  Hit* Ls = reinterpret_cast<Hit*>(smem);              // [W][k]
  Hit* sel = Ls + (size_t)W * p.k;                     // [k]
  int* cnt_s = reinterpret_cast<int*>(sel + p.k);      // [W]
  for (uint32_t qi = 0; qi < nqt; qi++) {
    for (uint32_t i = threadIdx.x; i < W * p.k; i += blockDim.x) {
      uint32_t w = i / p.k, e = i - w * p.k;
      if (e < (uint32_t)warp_cnt_s[w][qi]) Ls[i] = ld_hit(p.warp_lists + ((size_t)(blockIdx.x * W + w) * p.nq + (q0 + qi)) * p.k + e);
    }
    if (threadIdx.x < W) cnt_s[threadIdx.x] = warp_cnt_s[threadIdx.x][qi];
    __syncthreads();
    rank_merge_block(Ls, cnt_s, W, p.k, p.k, p.nearest, sel);
    __syncthreads();
    uint32_t total = 0;
    for (uint32_t w = 0; w < W; w++) total += (uint32_t)cnt_s[w];
    const uint32_t n_out = total < p.k ? total : p.k;
    Hit* out = p.cta_lists + ((size_t)blockIdx.x * p.nq + (q0 + qi)) * p.k;
    for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) st_hit(out + i, sel[i]);
    if (threadIdx.x == 0) p.cta_counts[(size_t)blockIdx.x * p.nq + (q0 + qi)] = (int)n_out;
    __syncthreads();
  }
  }  // query groups
}

int plan_flat_scan(int elem, uint32_t dim, uint32_t row_stride, uint32_t n_items, uint32_t nq, uint32_t k, int n_sms,
                   ScanPlan* plan) {
  if (k == 0 || k > 1024) return fail(COLTT_ERR_UNSUPPORTED, "top-k must be in [1, 1024]");
  if (nq == 0) return fail(COLTT_ERR_INVALID, "no queries");
  const uint32_t q_stride = (dim + 7) / 8 * 8;
  int qt = nq == 1 ? 1 : (nq <= 4 ? 4 : 8);
  // Two resident CTAs per SM with half-depth rings: measured on B200 (profiles/r2_k1_shapes.md) it lifts the one-query fp16
  // scan from 0.62 to 0.70 of the HBM roofline (the kernel is latency-bound at 2 warps per scheduler) but slows every other
  // shape (fp32 / f8 rows, 8-query groups: more per-CTA lists to merge, shallower rings), so it is used for that shape only.
  const int env_ctas = (elem == ELEM_F16 && nq == 1) ? 2 : 1;
  auto q_bytes = [&](int q) { return (size_t)q * q_stride * 4 + 32; };
  while (qt > 1 && q_bytes(qt) > 96 * 1024) qt = qt == 8 ? 4 : 1;
  if (q_bytes(qt) > 160 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "dim too large for the scan kernel");
  const size_t fixed = q_bytes(qt) + (size_t)kScanWarps * 8 * 8 + 128;
  const size_t half = (227 * 1024) / 2 - 1024 - 256;
  const size_t merge_need = ((size_t)kScanWarps * k + k) * sizeof(Hit) + kScanWarps * sizeof(int);
  const int ctas_per_sm = (env_ctas == 2 && fixed + 32 * 1024 <= half && merge_need <= half) ? 2 : 1;
  // static shared memory: counts (small) and, for the reference's f8 codec, the lane-private decode table
  const size_t static_smem = 256 + (elem == ELEM_F8C ? (size_t)kF8LutFloats * 4 : 0);
  const size_t budget = ctas_per_sm == 2 ? half : 227 * 1024 - static_smem;
  const size_t per_warp = (budget - fixed) / kScanWarps;
  const uint32_t row_up = (row_stride + 127) / 128 * 128;
  uint32_t cb = 0, stages = 0;
  for (uint32_t cand : {2048u, 1024u, 512u, 256u, 128u}) {
    uint32_t c = cand < row_up ? cand : row_up;
    uint32_t st = (uint32_t)(per_warp / ((size_t)kRowsPerWarp * (c + kRowPad)));
    if (st >= 3 || cand == 128u) {
      cb = c;
      stages = st > 8 ? 8 : st;
      break;
    }
  }
  if (stages < 1) return fail(COLTT_ERR_UNSUPPORTED, "no pipeline shape fits shared memory");
  const size_t merge_bytes = ((size_t)kScanWarps * k + k) * sizeof(Hit) + kScanWarps * sizeof(int);
  size_t smem = fixed + (size_t)kScanWarps * stages * kRowsPerWarp * (cb + kRowPad);
  if (smem < merge_bytes) smem = merge_bytes;
  if (smem + static_smem - 256 > 227 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "shared memory budget exceeded");
  const uint32_t n_groups = (n_items + kRowsPerWarp - 1) / kRowsPerWarp;
  uint32_t gx = (n_groups + kScanWarps - 1) / kScanWarps;
  if (gx > (uint32_t)(n_sms * ctas_per_sm)) gx = (uint32_t)(n_sms * ctas_per_sm);
  if (gx == 0) gx = 1;
  plan->grid_x = (int)gx;
  plan->grid_y = (int)((nq + qt - 1) / qt);
  plan->warps = kScanWarps;
  plan->qt = qt;
  plan->chunk_bytes = cb;
  plan->n_stages = stages;
  plan->smem_bytes = smem;
  plan->warp_list_bytes = (size_t)gx * kScanWarps * nq * k * sizeof(Hit);
  plan->cta_list_bytes = (size_t)gx * nq * k * sizeof(Hit);
  plan->cta_count_bytes = (size_t)gx * nq * sizeof(int);
  return COLTT_OK;
}

template <int ELEM, int METRIC>
static int launch_qt(const ScanParams& p, const ScanPlan& plan, cudaStream_t stream) {
  dim3 grid(plan.grid_x, plan.grid_y), block(kScanWarps * 32);
#define COLTT_LAUNCH(QT)                                                                                         \
  {                                                                                                              \
    auto kfn = flat_scan_kernel<ELEM, METRIC, QT>;                                                               \
    { int arc = kernel_attrs(kfn, plan.smem_bytes); if (arc) return arc; }    \
    kfn<<<grid, block, plan.smem_bytes, stream>>>(p);                                                            \
    count_launch();                                                                                              \
  }
  if (plan.qt == 1) COLTT_LAUNCH(1)
  else if (plan.qt == 4) COLTT_LAUNCH(4)
  else COLTT_LAUNCH(8)
#undef COLTT_LAUNCH
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

int launch_flat_scan(const ScanParams& p_in, const ScanPlan& plan, int elem, cudaStream_t stream) {
  ScanParams p = p_in;
  p.chunk_bytes = plan.chunk_bytes;
  p.n_stages = plan.n_stages;
  const bool cosine = p.metric == COLTT_COSINE;
  if (elem == ELEM_F32) return cosine ? launch_qt<ELEM_F32, COLTT_COSINE>(p, plan, stream) : launch_qt<ELEM_F32, COLTT_EUCLIDEAN>(p, plan, stream);
  if (elem == ELEM_F16) return cosine ? launch_qt<ELEM_F16, COLTT_COSINE>(p, plan, stream) : launch_qt<ELEM_F16, COLTT_EUCLIDEAN>(p, plan, stream);
  if (elem == ELEM_F8E) return cosine ? launch_qt<ELEM_F8E, COLTT_COSINE>(p, plan, stream) : launch_qt<ELEM_F8E, COLTT_EUCLIDEAN>(p, plan, stream);
  return cosine ? launch_qt<ELEM_F8C, COLTT_COSINE>(p, plan, stream) : launch_qt<ELEM_F8C, COLTT_EUCLIDEAN>(p, plan, stream);
}

}  // namespace coltt
