// hnsw_build.cu — bulk construction of a core/vectorindex HNSW graph on the GPU, and Hnsw.Commit.
//
// The reference builds the graph one Insert at a time (core/vectorindex/hnsw.go:104-167): greedy descent,
// searchLevel(efConstruction) on every level the new vertex lives on, connect to the M closest found
// (selectNeighbors, hnsw.go:391-397), add the back edges, and let every neighbour that now exceeds mMax
// (mMax0 on level 0) keep its closest mMax (pruneNeighbors, hnsw.go:449-474) — 2897 s for 1 M x 128 on the
// reference's own benchmark (benchmark/coltt_core.go:109).  A sequential replay of that on a GPU would run at
// the latency of one CTA.  The B200-native build computes the same graph in bulk, under one idealisation:
// searchLevel(efConstruction) is replaced by an EXACT search, i.e. the new vertex connects to its true M nearest
// among the vertices inserted before it.  With that, insertion order still matters but the graph has a closed form:
//   out(i)   = the M nearest members j < i of the level            (what Insert's search + selectNeighbors yield)
//   edges(v) = the closest mMax of  out(v)  U  { w : v in out(w) } (back edges; incremental pruneNeighbors always
//              keeps the closest, so the survivors are the closest mMax of everything ever added)
// and it is the graph the reference itself builds whenever its construction search is exhaustive (n up to about
// efConstruction) — tests/test_gpu_hnsw_build.py checks that case edge for edge against a CPU restatement of sequential
// Insert.  Steps:
//   1. levels: floor(-ln(U) * mL), mL = 1/ln(M) (hnsw.go:280-282, gomath/rand.go:42-44), U from a seeded
//      counter-based generator — or the caller's own levels, as Insert takes vertexLevel as an argument;
//      the first vertex gets level 0 and the entry point is the first vertex to reach the top level, as
//      sequential insertion leaves them (hnsw.go:109-118,160-163);
//   2. per level, members in insertion order, batches of 256: the batch is searched (k = M) against a temporary
//      fp16 FLAT shard holding the members BEFORE the batch — brute force on the tensor cores through the FAST
//      path (gemm_filter2.cu + rerank.cu), a triangular 1 M x 1 M x 768 / 2 at config-3 size — then joins the shard;
//   3. causal_select_kernel: per vertex, the shard's M hits plus its predecessors inside the batch are re-scored
//      with the exact fp32 arithmetic of the reference (AVX evaluation order, as hnsw.cu / flat_scan.cu) and the
//      M closest are kept, with the distances Commit stores per edge;
//   4. back edges + pruneNeighbors on the host (ties by id); edges are directed after pruning, as in Go.
// The result is a valid Hnsw in the reference's own terms — Commit() below writes the byte format
// hnsw_commit.go:69-162 defines, which the Go side can Load.  Beyond the exhaustive-search regime it differs from
// what Go would build only where Go's approximate construction search misses a true neighbour.
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>

#include "exact_math.cuh"
#include "hnsw.h"
#include "store.h"

namespace coltt {

int hnsw_load(const void* blob, size_t len, int device, Hnsw** out);

static constexpr uint32_t kNoNbr = 0xffffffffu;

// one warp per member row: contiguous fp32 copy of the level's rows (queries and ingest input of the temp shard)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* rows, uint32_t row_stride, const uint32_t* members, size_t m,
                                                          uint32_t dim, float* out) {
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= m) return;
  const float* src = reinterpret_cast<const float*>(rows + (size_t)members[w] * row_stride);
  float* dst = out + w * dim;
  for (uint32_t d = lane; d < dim; d += 32) dst[d] = src[d];
}

// Step 3.  One CTA per vertex of the batch: candidates = the prefix shard's hits (member indices < q0) and the
// batch members before it (q0 .. q0+q-1); exact fp32 distances, 64 candidate rows per pass (16 per warp, 2 lanes
// per row x 4 AVX-lane chains: pkg/distance/simd/cpp/avx.cpp:15-32,51-75); rank-select the M closest
// (NaN last, ties by insertion index).
static constexpr int kSelThreads = 128;
static constexpr uint32_t kSelBatch = 256;                      // vertices per batch
static constexpr uint32_t kSelMaxCand = kSelBatch + 32;         // hits (<= 24) + predecessors in the batch

template <int METRIC>
__global__ void __launch_bounds__(kSelThreads) causal_select_kernel(const uint8_t* rows, uint32_t row_stride, uint32_t dim, uint32_t q_stride,
                                                                    const float* norm2, const uint32_t* members, const Hit* hits,
                                                                    const int* hit_counts, uint32_t k_hits, uint32_t q0, uint32_t M,
                                                                    uint32_t* nbr, float* dist) {
  extern __shared__ __align__(16) float sel_smem[];
  float* q_s = sel_smem;                                        // [q_stride]
  float* c_dist = q_s + q_stride;                               // [kSelMaxCand]
  uint32_t* c_idx = reinterpret_cast<uint32_t*>(c_dist + kSelMaxCand);
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t self = q0 + q;
  const uint32_t v = members ? members[self] : self;
  const float* vrow = reinterpret_cast<const float*>(rows + (size_t)v * row_stride);
  for (uint32_t d = tid; d < q_stride; d += blockDim.x) q_s[d] = d < dim ? vrow[d] : 0.0f;
  const uint32_t nh = hits ? (uint32_t)hit_counts[q] : 0u;
  const uint32_t nc = nh + q;
  for (uint32_t e = tid; e < nc; e += blockDim.x) c_idx[e] = e < nh ? hits[(size_t)q * k_hits + e].slot : q0 + (e - nh);
  __syncthreads();
  const float qn = METRIC == COLTT_COSINE ? norm2[v] : 0.0f;
  const uint32_t r = lane_row16(lane), g = lane_half(lane);
  const uint32_t full8 = (dim / 8) * 8;
  for (uint32_t base = 0; base < nc; base += 64) {
    const uint32_t j = base + warp * 16 + r;
    const bool valid = j < nc;
    const uint32_t ci = valid ? c_idx[j] : self;
    const uint32_t u = members ? members[ci] : ci;
    const uint8_t* rowp = rows + (size_t)u * row_stride;
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
    for (uint32_t d = full8; d < dim; d++) {
      float rv = load1<ELEM_F32>(rowp, d, nullptr), qv = q_s[d];
      if (METRIC == COLTT_COSINE) tot = add_rn(tot, mul_rn(qv, rv));
      else { float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
    }
    if (valid && g == 0) c_dist[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, norm2[u]) : sqrt_via_f64(tot);
  }
  __syncthreads();
  uint32_t* out_n = nbr + (size_t)self * M;
  float* out_d = dist + (size_t)self * M;
  for (uint32_t e = tid; e < M; e += blockDim.x)
    if (e >= nc) out_n[e] = kNoNbr;
  for (uint32_t e = tid; e < nc; e += blockDim.x) {
    const float de = c_dist[e];
    const uint32_t ie = c_idx[e];
    const bool ne = de != de;
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nc; j++) {
      const float dj = c_dist[j];
      const uint32_t ij = c_idx[j];
      const bool nj = dj != dj;
      const bool before = (nj != ne) ? ne : ((!ne && dj != de) ? dj < de : ij < ie);   // j strictly ahead of e
      rank += before ? 1u : 0u;
    }
    if (rank < M) { out_n[rank] = ie; out_d[rank] = de; }
  }
}

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct LevelResult {
  std::vector<uint32_t> members;   // slots, ascending
  std::vector<uint32_t> nbr;       // [m][M] member indices or kNoNbr
  std::vector<float> dist;         // [m][M]
};

// back edges + pruneNeighbors for one level (hnsw.go:141-150,449-474), into lists[vbase[slot] + level]
static void assemble_level(const LevelResult& lr, uint32_t M, uint32_t cap, int level, const std::vector<uint32_t>& vbase,
                           const std::vector<uint64_t>& ids, std::vector<std::vector<HnswEdge>>& lists) {
  const size_t m = lr.members.size();
  std::vector<uint32_t> in_off(m + 1, 0);
  for (size_t i = 0; i < m; i++)
    for (uint32_t j = 0; j < M; j++) {
      const uint32_t u = lr.nbr[i * M + j];
      if (u != kNoNbr) in_off[u + 1]++;
    }
  for (size_t i = 0; i < m; i++) in_off[i + 1] += in_off[i];
  std::vector<uint32_t> in_src(in_off[m]);
  std::vector<float> in_dist(in_off[m]);
  {
    std::vector<uint32_t> fill(in_off.begin(), in_off.end() - 1);
    for (size_t i = 0; i < m; i++)
      for (uint32_t j = 0; j < M; j++) {
        const uint32_t u = lr.nbr[i * M + j];
        if (u == kNoNbr) continue;
        in_src[fill[u]] = (uint32_t)i;
        in_dist[fill[u]] = lr.dist[i * M + j];
        fill[u]++;
      }
  }
  auto work = [&](size_t lo, size_t hi) {
    std::vector<HnswEdge> cand;
    for (size_t i = lo; i < hi; i++) {
      cand.clear();
      auto add = [&](uint32_t mi, float d) {
        uint32_t bits;
        std::memcpy(&bits, &d, 4);
        cand.push_back(HnswEdge{ids[lr.members[mi]], lr.members[mi], bits});
      };
      for (uint32_t j = 0; j < M; j++)
        if (lr.nbr[i * M + j] != kNoNbr) add(lr.nbr[i * M + j], lr.dist[i * M + j]);
      for (uint32_t e = in_off[i]; e < in_off[i + 1]; e++) add(in_src[e], in_dist[e]);
      // the same pair reached from both sides carries the same distance bits (the arithmetic is symmetric)
      std::sort(cand.begin(), cand.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id < b.id; });
      cand.erase(std::unique(cand.begin(), cand.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id == b.id; }), cand.end());
      if (cand.size() > cap) {   // pruneNeighbors: keep the closest `cap` (NaN last, ties by id)
        auto key = [](const HnswEdge& e) { float f; std::memcpy(&f, &e.dist_bits, 4); return f; };
        std::sort(cand.begin(), cand.end(), [&](const HnswEdge& a, const HnswEdge& b) {
          const float da = key(a), db = key(b);
          const bool na = da != da, nb = db != db;
          if (na != nb) return nb;
          if (!na && da != db) return da < db;
          return a.id < b.id;
        });
        cand.resize(cap);
      }
      lists[vbase[lr.members[i]] + level] = cand;
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 32) nt = 32;
  if (m < 4096) nt = 1;
  std::vector<std::thread> th;
  const size_t per = (m + nt - 1) / nt;
  for (unsigned t = 0; t < nt; t++) {
    const size_t lo = t * per, hi = std::min(m, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
}

static int hnsw_build(const coltt_hnsw_build_cfg* cfg, const uint64_t* ids_in, const float* vecs, const int32_t* levels_in, size_t n_in,
                      Hnsw** out) {
  if (!cfg || !out) return fail(COLTT_ERR_INVALID, "null argument");
  if (cfg->dim == 0) return fail(COLTT_ERR_INVALID, "dim must be > 0");
  if (cfg->metric != COLTT_COSINE && cfg->metric != COLTT_EUCLIDEAN) return fail(COLTT_ERR_INVALID, "bad metric");
  if (n_in && (!ids_in || !vecs)) return fail(COLTT_ERR_INVALID, "null argument");
  if (n_in > 0xfffffff0ull) return fail(COLTT_ERR_UNSUPPORTED, "more than 2^32 vertices per GPU");
  const int32_t M = cfg->m > 0 ? cfg->m : 16;                  // hnsw_config.go:135-162 defaults
  if (M < 2 || M > 24) return fail(COLTT_ERR_UNSUPPORTED, "bulk build: m must be in [2, 24] (the batched FAST search returns at most 24 hits)");
  int rc = require_device(cfg->device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(cfg->device));
  const uint32_t n = (uint32_t)n_in, dim = cfg->dim;
  std::unique_ptr<Hnsw> h(new Hnsw());
  h->device = cfg->device;
  h->metric = cfg->metric;
  h->dim = dim;
  h->row_stride = (dim * 4 + 15) / 16 * 16;
  h->m = M; h->m_max = M; h->m_max0 = 2 * M;
  h->ef_default = cfg->ef > 0 ? cfg->ef : 20;
  h->ef_construction = cfg->ef_construction > 0 ? cfg->ef_construction : 200;
  h->search_algo = 0;   // HnswSearchSimple
  // levelMultiplier = 1/ln(m) through gomath.Log (float64 log -> f32), f32 divide (hnsw_config.go:155)
  const float level_mult = 1.0f / (float)std::log((double)(float)M);
  std::memcpy(&h->level_mult_bits, &level_mult, 4);
  cudaDeviceProp pr;
  COLTT_CUDA(cudaGetDeviceProperties(&pr, cfg->device));
  h->n_sms = pr.multiProcessorCount;
  COLTT_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  COLTT_CUDA(cudaMalloc((void**)&h->d_stats, 8 * sizeof(unsigned long long)));
  h->n = n;
  cudaStream_t st = h->stream;
  const double t0 = now_ms();

  // ---- vertices: ids must be unique (a second Insert of an id would replace the first in the Go map)
  std::vector<uint64_t> ids(ids_in, ids_in + n);
  {
    std::vector<uint64_t> sorted(ids);
    std::sort(sorted.begin(), sorted.end());
    if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) return fail(COLTT_ERR_INVALID, "bulk build: duplicate vertex id");
  }
  std::vector<int32_t> levels(n, 0);
  int32_t max_level = 0;
  uint32_t entry = 0;
  for (uint32_t i = 1; i < n; i++) {      // vertex 0 is created at level 0 and becomes the first entry point (hnsw.go:109-118)
    int32_t l;
    if (levels_in) {
      l = levels_in[i];
      if (l < 0 || l > 64) return fail(COLTT_ERR_INVALID, "bulk build: vertex level out of range");
    } else {
      const uint32_t r24 = (uint32_t)(splitmix64(cfg->seed + i) >> 40);
      const float u = (float)(r24 + 1) * (1.0f / 16777216.0f);   // (0, 1]
      const float lf = -(float)std::log((double)u) * level_mult; // gomath.RandomExponential, rand.go:42-44
      l = (int32_t)std::floor((double)lf);                       // gomath.Floor, hnsw.go:281
    }
    levels[i] = l;
    if (l > max_level) { max_level = l; entry = i; }             // strictly greater: hnsw.go:160-163
  }
  h->entry = entry;
  h->max_level = max_level;
  if ((rc = upload(&h->d_ids, ids)) || (rc = upload(&h->d_level, levels))) return rc;
  COLTT_CUDA(cudaMalloc((void**)&h->d_rows, std::max<size_t>(n, 1) * h->row_stride));
  COLTT_CUDA(cudaMalloc((void**)&h->d_norm2, std::max<size_t>(n, 1) * 4));
  std::vector<uint32_t> vbase(n + 1, 0);
  for (uint32_t v = 0; v < n; v++) vbase[v + 1] = vbase[v] + (uint32_t)levels[v] + 1;
  std::vector<std::vector<HnswEdge>> lists(vbase[n]);
  if (n == 0) {
    if ((rc = hnsw_install_graph(h.get(), vbase, lists))) return rc;
    *out = h.release();
    return COLTT_OK;
  }

  // ---- ingest: Normalize (cosine; hnsw.go:105-107) -> fp32 rows + ||row||^2 in the AVX lane order
  {
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(n, (256u << 20) / ((size_t)dim * 4)));
    DeviceBuf stage;
    PinnedBuf pin;
    if ((rc = stage.ensure(chunk * dim * 4)) || (rc = pin.ensure(chunk * dim * 4))) return rc;
    for (size_t base = 0; base < n; base += chunk) {
      const size_t c = std::min(chunk, n - base);
      std::memcpy(pin.p, vecs + base * dim, c * (size_t)dim * 4);
      COLTT_CUDA(cudaMemcpyAsync(stage.p, pin.p, c * (size_t)dim * 4, cudaMemcpyHostToDevice, st));
      PrepParams pp{};
      pp.in = (const float*)stage.p; pp.n = c; pp.in_stride = dim; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
      pp.normalize = cfg->metric == COLTT_COSINE;
      pp.rows_out = h->d_rows; pp.row_stride = h->row_stride; pp.slot_base = (uint32_t)base;
      pp.norm2_out = h->d_norm2; pp.norm2_by_slot = 1;
      rc = launch_prep_rows(pp, ELEM_F32, st);
      if (rc) return rc;
      COLTT_CUDA(cudaStreamSynchronize(st));   // the staging buffers are reused by the next chunk
    }
  }
  const double t1 = now_ms();
  double ms_knn = 0, ms_dist = 0, ms_host = 0;
  float max_abs = 0.0f;                        // Euclidean only: does the data fit the fp16 candidate shard?
  if (cfg->metric != COLTT_COSINE)
    for (size_t i = 0; i < n * (size_t)dim; i++) { const float a = std::fabs(vecs[i]); if (!(a <= max_abs)) max_abs = a == a ? a : INFINITY; }

  // ---- per level, top down
  const uint32_t q_stride = (dim + 7) / 8 * 8;
  DeviceBuf d_members, d_contig, d_nbr, d_dist, d_hits, d_counts;
  for (int level = max_level; level >= 0; level--) {
    LevelResult lr;
    for (uint32_t v = 0; v < n; v++)
      if (levels[v] >= level) lr.members.push_back(v);
    const size_t m = lr.members.size();
    if (m < 2) continue;                      // a lone vertex has nobody to connect to
    const double ta = now_ms();
    const bool identity = m == n;             // level 0: every vertex
    if ((rc = d_members.ensure(m * 4)) || (rc = d_nbr.ensure(m * (size_t)M * 4)) || (rc = d_dist.ensure(m * (size_t)M * 4))) return rc;
    COLTT_CUDA(cudaMemcpyAsync(d_members.p, lr.members.data(), m * 4, cudaMemcpyHostToDevice, st));
    const float* contig;
    if (identity && h->row_stride == dim * 4) {
      contig = reinterpret_cast<const float*>(h->d_rows);
    } else {
      if ((rc = d_contig.ensure(m * (size_t)dim * 4))) return rc;
      gather_rows_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(h->d_rows, h->row_stride, (const uint32_t*)d_members.p, m, dim, (float*)d_contig.p);
      count_launch();
      COLTT_CUDA(cudaGetLastError());
      contig = (const float*)d_contig.p;
    }
    // The members inserted so far, as a temporary FLAT shard (ids = member indices = insertion order) holding the SAME fp32
    // rows as the index (appended raw, not re-normalized).  Cosine: an fp32 shard, which COLTT_MATH_FAST serves through its
    // fp16 shadow with results bit-identical to the exact fp32 scan — the M hits ARE the true fp32 top-M, nothing is lost
    // to a coarser ranking.  Euclidean rows are not unit-norm: the fp32 shard has no tensor-core path, so an fp16 shard
    // ranks the candidates, over-fetched (M + 8, at most 24) so that the fp32 re-scoring below picks the true M nearest
    // unless more than 8 rows swap ranks at the cut; values outside the fp16 range fall back to the exact fp32 shard.
    const bool f16_shard = cfg->metric != COLTT_COSINE && max_abs <= 60000.0f;
    coltt_store_cfg scfg{};
    scfg.dim = dim; scfg.metric = cfg->metric; scfg.quant = f16_shard ? COLTT_QUANT_F16 : COLTT_QUANT_NONE; scfg.device = cfg->device; scfg.capacity_hint = m;
    Store* shard_raw = nullptr;
    rc = Store::create(&scfg, &shard_raw);
    if (rc) return rc;
    std::unique_ptr<Store> shard(shard_raw);
    shard->raw_queries = !f16_shard;          // queries are the index's own (already normalized) rows: score them as they are
    COLTT_CUDA(cudaStreamSynchronize(st));    // `contig` and the member list are complete before other streams read them
    {
      auto ctx = shard->acquire_ctx(nullptr);
      if (!ctx) return fail(COLTT_ERR_CUDA, "could not create a search context");
      cudaStream_t ss = ctx->stream;
      const size_t B = kSelBatch;
      if ((rc = d_hits.ensure(B * (size_t)(M + 8) * sizeof(Hit))) || (rc = d_counts.ensure(B * 4))) return rc;
      const size_t smem = ((size_t)q_stride + 2 * kSelMaxCand) * 4;
      if (smem > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "dim too large for the neighbour-selection kernel");
      const uint32_t* mem_arg = identity ? nullptr : (const uint32_t*)d_members.p;
      for (size_t q0 = 0; q0 < m && !rc; q0 += B) {
        const size_t nq = std::min(B, m - q0);
        const size_t want = f16_shard ? std::min<size_t>(24, (size_t)M + 8) : (size_t)M;
        const uint32_t k = (uint32_t)std::min<size_t>(want, q0);           // hits wanted from the members before the batch
        if (k) {
          rc = shard->search_enqueue(*ctx, ss, contig + q0 * dim, nq, (int)k, COLTT_SELECT_NEAREST, COLTT_MATH_FAST, nullptr, 0, (Hit*)d_hits.p,
                                     (int*)d_counts.p, false);
          if (rc) break;
        }
        const Hit* hit_arg = k ? (const Hit*)d_hits.p : nullptr;
        if (cfg->metric == COLTT_COSINE) {
          if ((rc = kernel_attrs(causal_select_kernel<COLTT_COSINE>, smem))) break;
          causal_select_kernel<COLTT_COSINE><<<(unsigned)nq, kSelThreads, smem, ss>>>(h->d_rows, h->row_stride, dim, q_stride, h->d_norm2, mem_arg, hit_arg,
                                                                                     (const int*)d_counts.p, k, (uint32_t)q0, (uint32_t)M,
                                                                                     (uint32_t*)d_nbr.p, (float*)d_dist.p);
        } else {
          if ((rc = kernel_attrs(causal_select_kernel<COLTT_EUCLIDEAN>, smem))) break;
          causal_select_kernel<COLTT_EUCLIDEAN><<<(unsigned)nq, kSelThreads, smem, ss>>>(h->d_rows, h->row_stride, dim, q_stride, h->d_norm2, mem_arg,
                                                                                        hit_arg, (const int*)d_counts.p, k, (uint32_t)q0, (uint32_t)M,
                                                                                        (uint32_t*)d_nbr.p, (float*)d_dist.p);
        }
        count_launch();
        // the batch joins the shard (disjoint rows: the search above keeps reading the prefix while this writes)
        rc = shard->append_dev(contig + q0 * dim, nq, dim, 0, /*raw=*/true);
      }
      cudaError_t e = cudaStreamSynchronize(ss);
      shard->release_ctx(std::move(ctx));
      if (rc) return rc;
      if (e != cudaSuccess) return fail(COLTT_ERR_CUDA, std::string("bulk build: neighbour search failed: ") + cudaGetErrorString(e));
      COLTT_CUDA(cudaGetLastError());
    }
    {
      unsigned long long fb = 0;
      COLTT_CUDA(cudaMemcpy(&fb, shard->d_stat, 8, cudaMemcpyDeviceToHost));
      h->build_fast_queries += shard->fast_queries;
      h->build_fast_fallbacks += fb;
    }
    shard.reset();
    const double tb = now_ms();
    lr.nbr.resize(m * (size_t)M);
    lr.dist.resize(m * (size_t)M);
    COLTT_CUDA(cudaMemcpyAsync(lr.nbr.data(), d_nbr.p, m * (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaMemcpyAsync(lr.dist.data(), d_dist.p, m * (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaStreamSynchronize(st));
    const double tc = now_ms();
    assemble_level(lr, (uint32_t)M, (uint32_t)(level == 0 ? h->m_max0 : h->m_max), level, vbase, ids, lists);
    const double td = now_ms();
    ms_knn += tb - ta; ms_dist += tc - tb; ms_host += td - tc;
  }
  const double t2 = now_ms();
  if ((rc = hnsw_install_graph(h.get(), vbase, lists))) return rc;
  ms_host += now_ms() - t2;
  h->build_ms[0] = t1 - t0; h->build_ms[1] = ms_knn; h->build_ms[2] = ms_dist; h->build_ms[3] = ms_host;
  *out = h.release();
  return COLTT_OK;
}

// ---- Hnsw.Commit(w, header=true): core/vectorindex/hnsw_commit.go:69-162, config per hnsw_config.go:179-201,
// vectors as big-endian float32, metadata count 0 (metadata stays with the Go side, metadata.go:31-41)
struct BlobW {
  uint8_t* p; size_t cap, pos = 0;
  void be(uint64_t v, int nb) {
    if (p && pos + nb <= cap)
      for (int i = 0; i < nb; i++) p[pos + i] = (uint8_t)(v >> (8 * (nb - 1 - i)));
    pos += nb;
  }
};

static int hnsw_commit(Hnsw* h, void* buf, size_t* len) {
  std::lock_guard<std::mutex> lk(h->mu);
  COLTT_CUDA(cudaSetDevice(h->device));
  const uint32_t n = h->n, dim = h->dim;
  std::vector<int32_t> levels(n);
  std::vector<uint64_t> ids(n);
  std::vector<uint32_t> vbase(n + 1, 0);
  if (n) {
    COLTT_CUDA(cudaMemcpy(levels.data(), h->d_level, n * 4, cudaMemcpyDeviceToHost));
    COLTT_CUDA(cudaMemcpy(ids.data(), h->d_ids, n * 8, cudaMemcpyDeviceToHost));
    COLTT_CUDA(cudaMemcpy(vbase.data(), h->d_vbase, (n + 1) * 4, cudaMemcpyDeviceToHost));
  }
  size_t need = 7 * 4 + 4 + 1;
  if (n) {
    need += 8 + 16 * 4 + (size_t)n * (8 + 4 + (size_t)dim * 4 + 2);
    need += (size_t)n * 8 + (size_t)vbase[n] * 4 + h->n_edges * 12;
  }
  if (!buf) { *len = need; return COLTT_OK; }
  if (*len < need) { *len = need; return fail(COLTT_ERR_INVALID, "commit buffer too small"); }
  BlobW w{(uint8_t*)buf, *len};
  w.be((uint32_t)h->search_algo, 4);
  w.be(h->level_mult_bits, 4);
  w.be((uint32_t)h->ef_default, 4);
  w.be((uint32_t)h->ef_construction, 4);
  w.be((uint32_t)h->m, 4); w.be((uint32_t)h->m_max, 4); w.be((uint32_t)h->m_max0, 4);
  w.be(dim, 4);
  w.be(h->metric == COLTT_COSINE ? 1 : 2, 1);   // distToDistIdx, hnsw_commit.go:36-46
  if (n) {
    std::vector<uint32_t> edge_off(vbase[n] + 1), edge_nbr(h->n_edges), edge_dist(h->n_edges);
    COLTT_CUDA(cudaMemcpy(edge_off.data(), h->d_edge_off, edge_off.size() * 4, cudaMemcpyDeviceToHost));
    if (h->n_edges) {
      COLTT_CUDA(cudaMemcpy(edge_nbr.data(), h->d_edge_nbr, h->n_edges * 4, cudaMemcpyDeviceToHost));
      COLTT_CUDA(cudaMemcpy(edge_dist.data(), h->d_edge_dist, h->n_edges * 4, cudaMemcpyDeviceToHost));
    }
    std::vector<std::vector<uint32_t>> shard_slots(16);   // VERTICES_MAP_SHARD_COUNT, hnsw.go:35
    for (uint32_t v = 0; v < n; v++) shard_slots[shard_vertex(ids[v], 16)].push_back(v);
    w.be(ids[h->entry], 8);
    std::vector<uint8_t> rowbuf(h->row_stride);
    for (int sh = 0; sh < 16; sh++) {
      w.be((uint32_t)shard_slots[sh].size(), 4);
      for (uint32_t v : shard_slots[sh]) {
        w.be(ids[v], 8);
        w.be((uint32_t)levels[v], 4);
        COLTT_CUDA(cudaMemcpy(rowbuf.data(), h->d_rows + (size_t)v * h->row_stride, (size_t)dim * 4, cudaMemcpyDeviceToHost));
        for (uint32_t d = 0; d < dim; d++) {
          uint32_t bits;
          std::memcpy(&bits, rowbuf.data() + (size_t)d * 4, 4);
          w.be(bits, 4);
        }
        w.be(0, 2);   // metadata count
      }
    }
    for (int sh = 0; sh < 16; sh++)
      for (uint32_t v : shard_slots[sh]) {
        w.be(ids[v], 8);
        for (int l = levels[v]; l >= 0; l--) {
          const uint32_t e0 = edge_off[vbase[v] + l], e1 = edge_off[vbase[v] + l + 1];
          w.be(e1 - e0, 4);
          for (uint32_t e = e0; e < e1; e++) { w.be(ids[edge_nbr[e]], 8); w.be(edge_dist[e], 4); }
        }
      }
  }
  if (w.pos != need) return fail(COLTT_ERR_FORMAT, "commit: size accounting mismatch");
  *len = need;
  return COLTT_OK;
}

}  // namespace coltt

using coltt::fail;
using coltt::Hnsw;

extern "C" {
COLTT_API int coltt_b200_hnsw_build(const coltt_hnsw_build_cfg* cfg, const uint64_t* ids, const float* vecs, const int32_t* levels, size_t n,
                                    coltt_hnsw** out) {
  Hnsw* h = nullptr;
  int rc = coltt::hnsw_build(cfg, ids, vecs, levels, n, &h);
  if (rc == COLTT_OK) *out = reinterpret_cast<coltt_hnsw*>(h);
  return rc;
}
COLTT_API int coltt_b200_hnsw_commit(coltt_hnsw* h, void* buf, size_t* len) {
  if (!h || !len) return fail(COLTT_ERR_INVALID, "null argument");
  return coltt::hnsw_commit(reinterpret_cast<Hnsw*>(h), buf, len);
}
COLTT_API int coltt_b200_hnsw_build_fast_stats(coltt_hnsw* h, uint64_t* out2) {
  if (!h || !out2) return fail(COLTT_ERR_INVALID, "null argument");
  out2[0] = reinterpret_cast<Hnsw*>(h)->build_fast_queries;
  out2[1] = reinterpret_cast<Hnsw*>(h)->build_fast_fallbacks;
  return COLTT_OK;
}
COLTT_API int coltt_b200_hnsw_build_stats(coltt_hnsw* h, double* ms4, uint64_t* n_edges, int32_t* max_level) {
  if (!h || !ms4) return fail(COLTT_ERR_INVALID, "null argument");
  Hnsw* x = reinterpret_cast<Hnsw*>(h);
  for (int i = 0; i < 4; i++) ms4[i] = x->build_ms[i];
  if (n_edges) *n_edges = x->n_edges;
  if (max_level) *max_level = x->max_level;
  return COLTT_OK;
}
}
