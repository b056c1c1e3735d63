// coltt_oracle.cpp — CPU restatement of the reference's ANN hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under coltt_b200/ may include, link or call
// this file.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's
// cpu_baseline / --impl reference legs (as the checker / the timed CPU arm).
//
// Parity pinning: the reference (sjy-dv/coltt @1624b2f) holds NO golden vectors or
// known-answer tests for FLAT search, codecs or top-K.  The
// distance arithmetic is pinned against the reference's own C++ source
// (pkg/distance/simd/cpp/avx.cpp) compiled unmodified into oracle/_ref/ and
// compared bit-for-bit in tests/test_oracle.py; everything else (Go control flow,
// codecs, heaps) is pinned by source restatement only, each function citing the
// reference file:line it follows.  PQ (BASELINE config 5): parity unpinned (no
// reference arithmetic exists).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; no -ffast-math — the whole
// point is bit-exact IEEE evaluation order).
//
// Determinism note: the reference iterates Go maps (random order).
// This restatement iterates shard 0..15 and, inside a shard, ascending id.  Output
// order inside a group of equal scores is normalised to ascending id.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------
// optional hook: the reference's own AVX kernels (oracle/_ref/libcoltt_ref_avx.so)
// installed at run time through orc_set_ref_kernels().  When set, the FLAT/HNSW
// restatements call the reference's compiled code for dot/norm/L2 instead of the
// scalar lane-order emulation below (they are bit-identical; tests check that).
// ---------------------------------------------------------------------------
typedef void (*ref_cos_fn)(size_t, float*, float*, float*, float*);
typedef void (*ref_l2_fn)(size_t, float*, float*, float*);
static ref_cos_fn g_ref_cos = nullptr;
static ref_l2_fn g_ref_l2 = nullptr;

ORC_API void orc_set_ref_kernels(void* cos_fn, void* l2_fn) {
  g_ref_cos = (ref_cos_fn)cos_fn;
  g_ref_l2 = (ref_l2_fn)l2_fn;
}

// ---------------------------------------------------------------------------
// Codecs
// ---------------------------------------------------------------------------

static inline uint32_t f32bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float bitsf32(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// pkg/compresshelper/float16.go:237-270 (f16bitsToF32bits); bf16.go:233-266 is the
// same function under another name.
ORC_API uint32_t orc_f16bits_to_f32bits(uint16_t in) {
  uint32_t sign = (uint32_t)(in & 0x8000) << 16;
  uint32_t exp = (uint32_t)(in & 0x7c00) >> 10;
  uint32_t coef = (uint32_t)(in & 0x03ff) << 13;
  if (exp == 0x1f) {
    if (coef == 0) return sign | 0x7f800000u | coef;
    return sign | 0x7fc00000u | coef;
  }
  if (exp == 0) {
    if (coef == 0) return sign;
    exp++;
    while ((coef & 0x7f800000u) == 0) { coef <<= 1; exp--; }
    coef &= 0x007fffffu;
  }
  return sign | ((exp + (0x7f - 0xf)) << 23) | coef;
}

// pkg/compresshelper/float16.go:274-321 (f32bitsToF16bits), RNE incl. subnormals.
ORC_API uint16_t orc_f32bits_to_f16bits(uint32_t u32) {
  uint32_t sign = u32 & 0x80000000u;
  uint32_t exp = u32 & 0x7f800000u;
  uint32_t coef = u32 & 0x007fffffu;
  if (exp == 0x7f800000u) {
    uint32_t nanBit = 0;
    if (coef != 0) nanBit = 0x0200u;
    return (uint16_t)((sign >> 16) | 0x7c00u | nanBit | (coef >> 13));
  }
  uint32_t halfSign = sign >> 16;
  int32_t unbiasedExp = (int32_t)(exp >> 23) - 127;
  int32_t halfExp = unbiasedExp + 15;
  if (halfExp >= 0x1f) return (uint16_t)(halfSign | 0x7c00u);
  if (halfExp <= 0) {
    if (14 - halfExp > 24) return (uint16_t)halfSign;
    uint32_t c = coef | 0x00800000u;
    uint32_t halfCoef = c >> (uint32_t)(14 - halfExp);
    uint32_t roundBit = 1u << (uint32_t)(13 - halfExp);
    if ((c & roundBit) != 0 && (c & (3 * roundBit - 1)) != 0) halfCoef++;
    return (uint16_t)(halfSign | halfCoef);
  }
  uint32_t uHalfExp = (uint32_t)halfExp << 10;
  uint32_t halfCoef = coef >> 13;
  uint32_t roundBit = 0x00001000u;
  if ((coef & roundBit) != 0 && (coef & (3 * roundBit - 1)) != 0)
    return (uint16_t)((halfSign | uHalfExp | halfCoef) + 1);
  return (uint16_t)(halfSign | uHalfExp | halfCoef);
}

// pkg/compresshelper/float8.go:233-266 (F8bitsToF32bits) — restated literally,
// including the half-edited masks: `(in&0x7c) >> 10` is always 0.
ORC_API uint32_t orc_f8bits_to_f32bits(uint8_t in) {
  uint32_t sign = (uint32_t)(in & 0x80) << 8;
  uint32_t exp = (uint32_t)(in & 0x7c) >> 10;
  uint32_t coef = (uint32_t)(in & 0x03) << 13;
  if (exp == 0x1f) {
    if (coef == 0) return sign | 0x7f800000u | coef;
    return sign | 0x7fc00000u | coef;
  }
  if (exp == 0) {
    if (coef == 0) return sign;
    exp++;
    while ((coef & 0x7f800000u) == 0) { coef <<= 1; exp--; }
    coef &= 0x007fffffu;
  }
  return sign | ((exp + (0x7f - 0xf)) << 23) | coef;
}

// pkg/compresshelper/float8.go:270-313 (f32bitsToF8bits) — literal, incl. the
// `sign := u32 & 0x800000` (bit 23) and the uint8 truncations.
ORC_API uint8_t orc_f32bits_to_f8bits(uint32_t u32) {
  uint32_t sign = u32 & 0x800000u;
  uint32_t exp = u32 & 0x7f800000u;
  uint32_t coef = u32 & 0x007fffffu;
  if (exp == 0x7f800000u) {
    uint32_t nanBit = 0;
    if (coef != 0) nanBit = 0x0200u;
    return (uint8_t)((sign >> 8) | 0x7cu | nanBit | (coef >> 13));
  }
  uint32_t halfSign = sign >> 8;
  int32_t unbiasedExp = (int32_t)(exp >> 23) - 127;
  int32_t halfExp = unbiasedExp + 15;
  if (halfExp >= 0x1f) return (uint8_t)(halfSign | 0x7cu);
  if (halfExp <= 0) {
    if (14 - halfExp > 24) return (uint8_t)halfSign;
    uint32_t c = coef | 0x00800000u;
    uint32_t halfCoef = c >> (uint32_t)(14 - halfExp);
    uint32_t roundBit = 1u << (uint32_t)(13 - halfExp);
    if ((c & roundBit) != 0 && (c & (3 * roundBit - 1)) != 0) halfCoef++;
    return (uint8_t)(halfSign | halfCoef);
  }
  uint32_t uHalfExp = (uint32_t)halfExp << 10;
  uint32_t halfCoef = coef >> 13;
  uint32_t roundBit = 0x00001000u;
  if ((coef & roundBit) != 0 && (coef & (3 * roundBit - 1)) != 0)
    return (uint8_t)((halfSign | uHalfExp | halfCoef) + 1);
  return (uint8_t)(halfSign | uHalfExp | halfCoef);
}

ORC_API void orc_f32_to_f16_array(const float* in, uint16_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = orc_f32bits_to_f16bits(f32bits(in[i]));
}
ORC_API void orc_f16_to_f32_array(const uint16_t* in, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = bitsf32(orc_f16bits_to_f32bits(in[i]));
}
ORC_API void orc_f32_to_f8_array(const float* in, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = orc_f32bits_to_f8bits(f32bits(in[i]));
}
ORC_API void orc_f8_to_f32_array(const uint8_t* in, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = bitsf32(orc_f8bits_to_f32bits(in[i]));
}

// ---------------------------------------------------------------------------
// F8_E4M3 — BUILDER-DEFINED extension, NOT reference arithmetic (the reference's f8
// codec is broken, so BASELINE config 4's performance mode is a real fp8 store).  PARITY UNPINNED against
// the reference by construction; the codec itself is pinned against an independent IEEE-style implementation
// (torch.float8_e4m3fn) in tests/test_oracle.py.  Format: OCP FP8 E4M3 ("fn": no infinities, 0x7f = NaN,
// max 448), round-to-nearest-even, saturating.  A vector v is lowered as
//     s     = 2^clamp(floor(log2(max|v_i|)) - 7, -40, 40)        (1.0 when max|v_i| is 0 or not finite)
//     c_i   = e4m3_rne_sat(v_i / s)                               (so max|v_i|/s lies in [128, 256))
//     x_i   = s * e4m3_decode(c_i)                                (the dequantized value; exact: s is a power of two)
// and, as in every quantized store of the reference ({f16,bf16,f8}_quantization.go:33-43), Similarity
// dequantizes BOTH operands (the query is lowered too, f8_vectorstore.go:136-139) and calls dist.Distance.
// ---------------------------------------------------------------------------
ORC_API float orc_e4m3_decode(uint8_t c) {
  const int e = (c >> 3) & 15, m = c & 7;
  float v;
  if (e == 15 && m == 7) v = std::numeric_limits<float>::quiet_NaN();
  else if (e == 0) v = std::ldexp((float)m, -9);
  else v = std::ldexp((float)(8 + m), e - 10);
  return (c & 0x80) ? -v : v;
}
ORC_API uint8_t orc_e4m3_encode(float x) {
  const uint8_t sign = std::signbit(x) ? 0x80 : 0x00;
  const float a = std::fabs(x);
  if (a != a) return (uint8_t)0x7f;                      // one canonical NaN code (the sign of a NaN is not part of the contract)
  if (a > 448.0f) return (uint8_t)(sign | 0x7e);        // saturate (values in (448, 464) would round down to 448 anyway)
  if (a < 0.015625f) {                                   // below 2^-6: subnormal grid of 2^-9
    const int q = (int)std::nearbyint(std::ldexp(a, 9)); // exact scaling, RNE (default rounding mode)
    return (uint8_t)(sign | q);                          // q == 8 is the smallest normal, 0x08
  }
  int e;
  const float m = std::frexp(a, &e);                     // a = m * 2^e, m in [0.5, 1)
  int q = (int)std::nearbyint(std::ldexp(m, 4));         // significand on a grid of 1/8: q in [8, 16]
  e -= 1;
  if (q == 16) { q = 8; e += 1; }
  return (uint8_t)(sign | ((e + 7) << 3) | (q - 8));
}
ORC_API float orc_e4m3_scale(const float* v, size_t n) {
  float mx = 0.0f;
  bool bad = false;
  for (size_t i = 0; i < n; i++) { const float a = std::fabs(v[i]); if (!(a <= std::numeric_limits<float>::max())) bad = true; else if (a > mx) mx = a; }
  if (bad || mx == 0.0f) return 1.0f;
  int e;
  std::frexp(mx, &e);                                    // mx in [2^(e-1), 2^e)
  int es = (e - 1) - 7;
  if (es < -40) es = -40;
  if (es > 40) es = 40;
  return std::ldexp(1.0f, es);
}
ORC_API float orc_f32_to_e4m3_array(const float* in, uint8_t* out, size_t n) {
  const float s = orc_e4m3_scale(in, n);
  const float inv = 1.0f / s;                            // exact: power of two
  for (size_t i = 0; i < n; i++) out[i] = orc_e4m3_encode(in[i] * inv);
  return s;
}
ORC_API void orc_e4m3_to_f32_array(const uint8_t* in, float scale, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = scale * orc_e4m3_decode(in[i]);
}

// ---------------------------------------------------------------------------
// Normalize — edge/vectorstore.go:173-189 == core/vectorindex/metadata.go:107-123.
// Sequential f32 sum (Go amd64 does not fuse x*y+z), float32(sqrt(float64)), f32 divide.
// ---------------------------------------------------------------------------
ORC_API void orc_normalize(const float* v, size_t d, float* out) {
  volatile float norm = 0.0f;  // volatile: forbid reassociation/vectorised reduction
  for (size_t i = 0; i < d; i++) {
    float sq = v[i] * v[i];
    norm = norm + sq;
  }
  float n = norm;
  if (n == 0.0f) {
    for (size_t i = 0; i < d; i++) out[i] = 0.0f;
    return;
  }
  n = (float)std::sqrt((double)n);
  for (size_t i = 0; i < d; i++) out[i] = v[i] / n;
}

// ---------------------------------------------------------------------------
// Distance kernels: scalar emulation of pkg/distance/simd/cpp/avx.cpp evaluation
// order (== simd/avx/AVX_amd64.s).  8 lane accumulators, element i -> lane i%8 in
// increasing i, unfused mul then add, reduction ((l0+l1)+(l2+l3))+((l4+l5)+(l6+l7))
// (avx.cpp:3-8: two hadd_ps then [0]+[4]), then the scalar tail added in order.
// ---------------------------------------------------------------------------
static inline float sum_vector8(const float l[8]) {  // avx.cpp:3-8
  float a = (l[0] + l[1]) + (l[2] + l[3]);
  float b = (l[4] + l[5]) + (l[6] + l[7]);
  return a + b;
}

// avx.cpp:51-75 cosine_similarity_dot_norm
ORC_API void orc_cosine_dot_norm(size_t len, const float* a, const float* b, float* result_dot,
                                 float* result_norm_squared) {
  float dot[8] = {0}, na[8] = {0}, nb[8] = {0};
  size_t full = (len / 8) * 8;
  for (size_t i = 0; i < full; i += 8) {
    for (int j = 0; j < 8; j++) {
      float v1 = a[i + j], v2 = b[i + j];
      float p = v1 * v2; dot[j] = dot[j] + p;
      float pa = v1 * v1; na[j] = na[j] + pa;
      float pb = v2 * v2; nb[j] = nb[j] + pb;
    }
  }
  float d = sum_vector8(dot), sa = sum_vector8(na), sb = sum_vector8(nb);
  for (size_t i = full; i < len; i++) {
    float p = a[i] * b[i]; d = d + p;
    float pa = a[i] * a[i]; sa = sa + pa;
    float pb = b[i] * b[i]; sb = sb + pb;
  }
  *result_dot = d;
  *result_norm_squared = sa * sb;
}

// avx.cpp:15-32 euclidean_distance_squared
ORC_API void orc_l2sq(size_t len, const float* a, const float* b, float* result) {
  float acc[8] = {0};
  size_t full = (len / 8) * 8;
  for (size_t i = 0; i < full; i += 8) {
    for (int j = 0; j < 8; j++) {
      float df = a[i + j] - b[i + j];
      float sq = df * df;
      acc[j] = acc[j] + sq;
    }
  }
  float r = sum_vector8(acc);
  for (size_t i = full; i < len; i++) {
    float df = a[i] - b[i];
    float sq = df * df;
    r = r + sq;
  }
  *result = r;
}

// The ||x||^2 the cosine kernel computes for one operand alone (used by the CUDA path,
// which precomputes it per row at ingest).  Same lanes/tree as above.
ORC_API float orc_norm2_avx_order(size_t len, const float* a) {
  float na[8] = {0};
  size_t full = (len / 8) * 8;
  for (size_t i = 0; i < full; i += 8)
    for (int j = 0; j < 8; j++) { float p = a[i + j] * a[i + j]; na[j] = na[j] + p; }
  float s = sum_vector8(na);
  for (size_t i = full; i < len; i++) { float p = a[i] * a[i]; s = s + p; }
  return s;
}

// The reference's _mm256_load_ps needs 32-byte alignment (avx.cpp:19-20); callers of the
// reference hook pass aligned scratch copies.
struct AlignedBuf {
  float* p = nullptr; size_t cap = 0;
  float* get(size_t n) {
    if (n > cap) { std::free(p); p = (float*)std::aligned_alloc(32, ((n * 4 + 31) / 32) * 32); cap = n; }
    return p;
  }
  ~AlignedBuf() { std::free(p); }
};

// simd/avx/AVX_amd64.go:44-52 CosineDistance + space.go:93-95 Cosine.Distance
// (gomath.Abs, math.go:35-37).
static inline float cosine_distance_impl(size_t len, const float* a, const float* b, bool aligned) {
  float dot, n2;
  if (g_ref_cos && aligned) g_ref_cos(len, (float*)a, (float*)b, &dot, &n2);
  else orc_cosine_dot_norm(len, a, b, &dot, &n2);
  float q = dot / (float)std::sqrt((double)n2);
  float r = 1.0f - q;
  return (float)std::fabs((double)r);
}
// simd/avx/AVX_amd64.go:26-32 EuclideanDistance (sqrt via float64) + space.go:61-63
static inline float euclid_distance_impl(size_t len, const float* a, const float* b, bool aligned) {
  float r;
  if (g_ref_l2 && aligned) g_ref_l2(len, (float*)a, (float*)b, &r);
  else orc_l2sq(len, a, b, &r);
  return (float)std::sqrt((double)r);
}
static inline bool is_aligned32(const void* p) { return ((uintptr_t)p & 31) == 0; }

ORC_API float orc_cosine_distance(size_t len, const float* a, const float* b) {
  return cosine_distance_impl(len, a, b, is_aligned32(a) && is_aligned32(b));
}
ORC_API float orc_euclidean_distance(size_t len, const float* a, const float* b) {
  return euclid_distance_impl(len, a, b, is_aligned32(a) && is_aligned32(b));
}

// edge/edge_helper.go:143-148 scoreHelper
ORC_API float orc_score_helper(float score, int metric) {
  if (metric == 0) return ((2 - score) / 2) * 100;
  return (float)std::max(0.0, (double)(100 - score));
}

// experimental/multi_vector_vertex.go:60-67,85-137 MultiVertexSearch (CFLAT), restated over dense arrays:
// fields[f] is the [n][dim] matrix of field f as handed to ChangedVertex (normalized here for cosine, :64-66);
// q_field[j] / queries[j] / ratios[j] describe the INCLUDED query vectors in request order.  Per vertex
//   score = 0; score += scoreHelper(Distance(vertex[field_j], Normalize(query_j))) * (float32(ratio_j) / 100)
// (:108-116, float32, unfused), then the topK largest scores, descending (multi_priority_queue.go:47-75).  Among equal
// scores the reference's order depends on Go's heap and map iteration; this checker uses score descending, NaN first,
// then id descending (the reverse of the total order T used for FLAT).
ORC_API int orc_multi_search(uint32_t dim, int metric, size_t n, const uint64_t* ids, int n_fields, const float* const* fields,
                             int n_inc, const int* q_field, const float* const* queries, const int* ratios, int k,
                             uint64_t* out_ids, float* out_scores) {
  if (n_inc <= 0 || k <= 0) return 0;
  std::vector<AlignedBuf> stored(n_fields);
  std::vector<float*> sp(n_fields);
  const size_t dpad = ((size_t)dim + 7) / 8 * 8;
  for (int f = 0; f < n_fields; f++) {
    sp[f] = stored[f].get(std::max<size_t>(n, 1) * dpad);
    for (size_t r = 0; r < n; r++) {
      float* dst = sp[f] + r * dpad;
      if (metric == 0 /* cosine */) orc_normalize(fields[f] + r * (size_t)dim, dim, dst);
      else std::memcpy(dst, fields[f] + r * (size_t)dim, 4 * (size_t)dim);
    }
  }
  std::vector<AlignedBuf> qb(n_inc);
  std::vector<float*> qp(n_inc);
  for (int j = 0; j < n_inc; j++) {
    qp[j] = qb[j].get(dpad);
    if (metric == 0 /* cosine */) orc_normalize(queries[j], dim, qp[j]);          // :97-101
    else std::memcpy(qp[j], queries[j], 4 * (size_t)dim);
  }
  struct MScore { float priority; uint64_t id; };
  std::vector<MScore> all(n);
  for (size_t r = 0; r < n; r++) {
    float score = 0.0f;
    for (int j = 0; j < n_inc; j++) {
      const float* v = sp[q_field[j]] + r * dpad;
      const float sim = metric == 0 /* cosine */ ? cosine_distance_impl(dim, v, qp[j], true) : euclid_distance_impl(dim, v, qp[j], true);
      const float w = (float)(uint32_t)ratios[j] / 100;
      const float term = orc_score_helper(sim, metric) * w;
      score = score + term;
    }
    all[r] = MScore{score, ids[r]};
  }
  auto before = [](const MScore& a, const MScore& b) {                      // reverse of T
    const bool an = a.priority != a.priority, bn = b.priority != b.priority;
    if (an != bn) return an;
    if (!an && a.priority != b.priority) return a.priority > b.priority;
    return a.id > b.id;
  };
  const size_t kk = std::min<size_t>((size_t)k, n);
  std::partial_sort(all.begin(), all.begin() + kk, all.end(), before);
  for (size_t i = 0; i < kk; i++) { out_ids[i] = all[i].id; out_scores[i] = all[i].priority; }
  return (int)kk;
}

// pkg/sharding/shard.go:34-41 ShardVertex: FNV-1a 64 over the LE bytes of id, mod c.
ORC_API uint64_t orc_shard_vertex(uint64_t x, uint64_t c) {
  uint64_t h = 14695981039346656037ull;
  for (int i = 0; i < 8; i++) {
    h ^= (uint64_t)((x >> (8 * i)) & 0xff);
    h *= 1099511628211ull;
  }
  return h % c;
}

// ---------------------------------------------------------------------------
// Go container/heap (Go 1.23 standard library, src/container/heap/heap.go: Push =
// append+up, Pop = swap(0,n-1)+down+remove last), instantiated the way
// edge/priorityqueue/priority_queue.go:160-198 and
// core/vectorindex/priority_queue.go do: min-queue Less = a<b, max-queue Less = a>b.
// ---------------------------------------------------------------------------
struct HeapItem { float priority; uint64_t id; };

struct GoHeap {
  std::vector<HeapItem> q;
  bool is_max;
  explicit GoHeap(bool mx) : is_max(mx) {}
  inline bool less(size_t i, size_t j) const {
    return is_max ? (q[i].priority > q[j].priority) : (q[i].priority < q[j].priority);
  }
  void up(size_t j) {
    for (;;) {
      if (j == 0) break;
      size_t i = (j - 1) / 2;
      if (i == j || !less(j, i)) break;
      std::swap(q[i], q[j]);
      j = i;
    }
  }
  void down(size_t i0, size_t n) {
    size_t i = i0;
    for (;;) {
      size_t j1 = 2 * i + 1;
      if (j1 >= n) break;
      size_t j = j1;
      size_t j2 = j1 + 1;
      if (j2 < n && less(j2, j1)) j = j2;
      if (!less(j, i)) break;
      std::swap(q[i], q[j]);
      i = j;
    }
  }
  void push(HeapItem it) { q.push_back(it); up(q.size() - 1); }
  HeapItem pop() {
    size_t n = q.size() - 1;
    std::swap(q[0], q[n]);
    down(0, n);
    HeapItem it = q.back();
    q.pop_back();
    return it;
  }
  size_t len() const { return q.size(); }
  const HeapItem& peek() const { return q[0]; }  // priority_queue.go Peek = ToSlice()[0]
};

// edge/priority_queue.go:27-75: bounded queue over a MIN heap; Add = Push, then Pop the
// minimum when Len > maxSize (=> keeps the K LARGEST scores); ToSlice
// sorts ascending by Score.  select_mode 1 (NEAREST, builder extension) flips to a max
// heap so the K smallest are kept.
struct EdgePQ {
  GoHeap h;
  size_t max_size;
  EdgePQ(size_t k, int select_mode) : h(select_mode == 1), max_size(k) {}
  void add(float score, uint64_t id) {
    h.push({score, id});
    if (h.len() > max_size) h.pop();
  }
  std::vector<HeapItem> to_slice() const {
    std::vector<HeapItem> r = h.q;
    // sort.Slice ascending by Score (priority_queue.go:65-67); ties normalised to id order.
    std::stable_sort(r.begin(), r.end(), [](const HeapItem& a, const HeapItem& b) {
      if (a.priority < b.priority) return true;
      if (b.priority < a.priority) return false;
      return a.id < b.id;
    });
    return r;
  }
};

// ---------------------------------------------------------------------------
// edge FLAT store restatement: {none,f16,bf16,f8}_vectorstore.go
// ---------------------------------------------------------------------------
enum { Q_NONE = 0, Q_F16 = 1, Q_F8 = 2, Q_BF16 = 3, Q_F8E = 16 };  // idl/proto/v4/edge.proto:75-80; 16 = builder-defined E4M3 (see codecs)
enum { M_COSINE = 0, M_EUCLID = 1 };                  // edge.proto:69-72
static const int kShards = 16;                        // edge/constants.go:49

struct OrcStore {
  uint32_t dim; int metric; int quant;
  // per shard: id -> slot in the shard-local contiguous arrays (ordered => ascending id)
  std::map<uint64_t, size_t> index[kShards];
  std::vector<float> f32[kShards];
  std::vector<uint16_t> f16[kShards];
  std::vector<uint8_t> f8[kShards];
  std::vector<float> scale[kShards];   // Q_F8E: per-row power-of-two scale
  std::vector<size_t> free_slots[kShards];
};

ORC_API OrcStore* orc_store_create(uint32_t dim, int metric, int quant) {
  if (!((quant >= 0 && quant <= 3) || quant == Q_F8E) || metric < 0 || metric > 1 || dim == 0) return nullptr;
  OrcStore* s = new OrcStore();
  s->dim = dim; s->metric = metric; s->quant = quant;
  return s;
}
ORC_API void orc_store_destroy(OrcStore* s) { delete s; }
ORC_API uint64_t orc_store_size(OrcStore* s) {
  uint64_t n = 0; for (int i = 0; i < kShards; i++) n += s->index[i].size(); return n;
}

// ChangedVertex: none_vectorstore.go:66-103 / bf16_vectorstore.go:66-106 —
// normalize when cosine, Lower, store under ShardVertex(id,16).  (Metadata, inverted
// index and the primary-key lookup belong to the Go caller, out of scope.)
ORC_API int orc_store_upsert(OrcStore* s, const uint64_t* ids, const float* vecs, size_t n) {
  std::vector<float> tmp(s->dim);
  for (size_t r = 0; r < n; r++) {
    const float* v = vecs + r * (size_t)s->dim;
    if (s->metric == M_COSINE) { orc_normalize(v, s->dim, tmp.data()); v = tmp.data(); }
    int sh = (int)orc_shard_vertex(ids[r], kShards);
    size_t slot;
    auto it = s->index[sh].find(ids[r]);
    if (it != s->index[sh].end()) slot = it->second;
    else {
      if (!s->free_slots[sh].empty()) { slot = s->free_slots[sh].back(); s->free_slots[sh].pop_back(); }
      else {
        slot = s->index[sh].size() + s->free_slots[sh].size();
        size_t need = (slot + 1) * (size_t)s->dim;
        if (s->quant == Q_NONE) { if (s->f32[sh].size() < need) s->f32[sh].resize(need); }
        else if (s->quant == Q_F8) { if (s->f8[sh].size() < need) s->f8[sh].resize(need); }
        else if (s->quant == Q_F8E) { if (s->f8[sh].size() < need) s->f8[sh].resize(need); if (s->scale[sh].size() < slot + 1) s->scale[sh].resize(slot + 1); }
        else { if (s->f16[sh].size() < need) s->f16[sh].resize(need); }
      }
      s->index[sh][ids[r]] = slot;
    }
    size_t off = slot * (size_t)s->dim;
    if (s->quant == Q_NONE) std::memcpy(&s->f32[sh][off], v, 4 * (size_t)s->dim);
    else if (s->quant == Q_F8) orc_f32_to_f8_array(v, &s->f8[sh][off], s->dim);     // f8_quantization.go:45-51
    else if (s->quant == Q_F8E) s->scale[sh][slot] = orc_f32_to_e4m3_array(v, &s->f8[sh][off], s->dim);
    else orc_f32_to_f16_array(v, &s->f16[sh][off], s->dim);                         // f16/bf16_quantization.go Lower
  }
  return 0;
}

// RemoveVertex (none_vectorstore.go:105-127) after the Go side resolved the filter to ids.
ORC_API int orc_store_remove(OrcStore* s, const uint64_t* ids, size_t n) {
  for (size_t r = 0; r < n; r++) {
    int sh = (int)orc_shard_vertex(ids[r], kShards);
    auto it = s->index[sh].find(ids[r]);
    if (it == s->index[sh].end()) continue;
    s->free_slots[sh].push_back(it->second);
    s->index[sh].erase(it);
  }
  return 0;
}

struct QueryCtx {
  std::vector<float> q32;       // normalized (cosine) query, fp32 path
  std::vector<uint16_t> q16;    // Lower(query) for f16/bf16 stores
  std::vector<uint8_t> q8;      // Lower(query) for f8 store
  float q_scale = 1.0f;         // Q_F8E
};

static void prep_query(const OrcStore* s, const float* query, QueryCtx& c) {
  c.q32.resize(s->dim);
  if (s->metric == M_COSINE) orc_normalize(query, s->dim, c.q32.data());   // *_vectorstore.go:131-134
  else std::memcpy(c.q32.data(), query, 4 * (size_t)s->dim);
  if (s->quant == Q_F16 || s->quant == Q_BF16) { c.q16.resize(s->dim); orc_f32_to_f16_array(c.q32.data(), c.q16.data(), s->dim); }
  if (s->quant == Q_F8) { c.q8.resize(s->dim); orc_f32_to_f8_array(c.q32.data(), c.q8.data(), s->dim); }
  if (s->quant == Q_F8E) { c.q8.resize(s->dim); c.q_scale = orc_f32_to_e4m3_array(c.q32.data(), c.q8.data(), s->dim); }
}

// Quantization.Similarity: quantization.go:43-45 (none) and
// {f16,bf16,f8}_quantization.go:33-43: dequantize BOTH operands into fresh buffers on every
// call, then dist.Distance(bufx=query, bufy=row).
struct SimScratch { AlignedBuf bx, by; };
static inline float similarity(const OrcStore* s, const QueryCtx& c, int sh, size_t slot, SimScratch& sc) {
  size_t d = s->dim, off = slot * d;
  const float *x, *y;
  if (s->quant == Q_NONE) {
    if (g_ref_cos) {  // reference kernels need 32B-aligned inputs: copy (bit-identical values)
      float* bx = sc.bx.get(d); float* by = sc.by.get(d);
      std::memcpy(bx, c.q32.data(), 4 * d); std::memcpy(by, &s->f32[sh][off], 4 * d);
      x = bx; y = by;
    } else { x = c.q32.data(); y = &s->f32[sh][off]; }
  } else {
    float* bx = sc.bx.get(d); float* by = sc.by.get(d);
    if (s->quant == Q_F8) {
      for (size_t i = 0; i < d; i++) { bx[i] = bitsf32(orc_f8bits_to_f32bits(c.q8[i])); by[i] = bitsf32(orc_f8bits_to_f32bits(s->f8[sh][off + i])); }
    } else if (s->quant == Q_F8E) {
      orc_e4m3_to_f32_array(c.q8.data(), c.q_scale, bx, d);
      orc_e4m3_to_f32_array(&s->f8[sh][off], s->scale[sh][slot], by, d);
    } else {
      for (size_t i = 0; i < d; i++) { bx[i] = bitsf32(orc_f16bits_to_f32bits(c.q16[i])); by[i] = bitsf32(orc_f16bits_to_f32bits(s->f16[sh][off + i])); }
    }
    x = bx; y = by;
  }
  bool al = is_aligned32(x) && is_aligned32(y);
  return s->metric == M_COSINE ? cosine_distance_impl(d, x, y, al) : euclid_distance_impl(d, x, y, al);
}

static int emit(const std::vector<HeapItem>& r, uint64_t* out_ids, float* out_scores) {
  for (size_t i = 0; i < r.size(); i++) { out_ids[i] = r[i].id; out_scores[i] = r[i].priority; }
  return (int)r.size();
}

// VertexSearch: none_vectorstore.go:129-180 (and twins).  high_cpu=0: one PQ over shards
// 0..15 serially; high_cpu=1: 16 shard-local PQs (goroutines; here n_threads workers), each
// ToSlice()d, then re-Added into the global PQ in shard order (:173-178).
// select_mode 0 = COLTT_COMPAT (literal reference semantics, K largest), 1 = NEAREST.
ORC_API int orc_store_search(OrcStore* s, const float* query, int top_k, int high_cpu, int select_mode,
                             int n_threads, uint64_t* out_ids, float* out_scores) {
  if (top_k <= 0) return 0;
  QueryCtx c; prep_query(s, query, c);
  EdgePQ pq((size_t)top_k, select_mode);
  if (!high_cpu) {
    SimScratch sc;
    for (int sh = 0; sh < kShards; sh++)
      for (auto& kv : s->index[sh]) pq.add(similarity(s, c, sh, kv.second, sc), kv.first);
  } else {
    std::vector<std::vector<HeapItem>> results(kShards);
    auto work = [&](int sh) {
      SimScratch sc;
      EdgePQ local((size_t)top_k, select_mode);
      for (auto& kv : s->index[sh]) local.add(similarity(s, c, sh, kv.second, sc), kv.first);
      results[sh] = local.to_slice();
    };
    int nt = std::max(1, std::min(n_threads, kShards));
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() { for (;;) { int sh = next.fetch_add(1); if (sh >= kShards) break; work(sh); } });
    for (auto& t : th) t.join();
    for (int sh = 0; sh < kShards; sh++) for (auto& it : results[sh]) pq.add(it.priority, it.id);
  }
  return emit(pq.to_slice(), out_ids, out_scores);
}

// FilterableVertexSearch: none_vectorstore.go:182-253 with the candidate id list the
// inverted index would return (pkg/inverted/search.go:113-119): bucket by ShardVertex,
// scan candidates in list order per shard, skip ids not present.
ORC_API int orc_store_search_subset(OrcStore* s, const float* query, const uint64_t* cand, size_t n_cand,
                                    int top_k, int high_cpu, int select_mode, uint64_t* out_ids, float* out_scores) {
  if (top_k <= 0) return 0;
  QueryCtx c; prep_query(s, query, c);
  std::vector<uint64_t> buckets[kShards];
  for (size_t i = 0; i < n_cand; i++) buckets[orc_shard_vertex(cand[i], kShards)].push_back(cand[i]);
  EdgePQ pq((size_t)top_k, select_mode);
  SimScratch sc;
  if (!high_cpu) {
    for (int sh = 0; sh < kShards; sh++)
      for (uint64_t id : buckets[sh]) {
        auto it = s->index[sh].find(id);
        if (it != s->index[sh].end()) pq.add(similarity(s, c, sh, it->second, sc), id);
      }
  } else {
    for (int sh = 0; sh < kShards; sh++) {
      EdgePQ local((size_t)top_k, select_mode);
      for (uint64_t id : buckets[sh]) {
        auto it = s->index[sh].find(id);
        if (it != s->index[sh].end()) local.add(similarity(s, c, sh, it->second, sc), id);
      }
      for (auto& item : local.to_slice()) pq.add(item.priority, item.id);
    }
  }
  return emit(pq.to_slice(), out_ids, out_scores);
}

// Total-order selector (builder-defined tie rule: a window of the (score,id)-sorted order).  Scores are computed exactly as above; only the
// selection is order-independent.  This is what the CUDA path implements; on tie-free
// inputs it equals orc_store_search (tests assert that).
ORC_API int orc_store_search_total_order(OrcStore* s, const float* query, const uint64_t* cand, size_t n_cand,
                                         int use_subset, int top_k, int select_mode, int n_threads,
                                         uint64_t* out_ids, float* out_scores) {
  if (top_k <= 0) return 0;
  QueryCtx c; prep_query(s, query, c);
  std::vector<HeapItem> all;
  if (use_subset) {
    SimScratch sc;
    std::unordered_set<uint64_t> seen;
    for (size_t i = 0; i < n_cand; i++) {
      if (!seen.insert(cand[i]).second) continue;
      int sh = (int)orc_shard_vertex(cand[i], kShards);
      auto it = s->index[sh].find(cand[i]);
      if (it != s->index[sh].end()) all.push_back({similarity(s, c, sh, it->second, sc), cand[i]});
    }
  } else {
    std::vector<std::vector<HeapItem>> part(kShards);
    int nt = std::max(1, std::min(n_threads, kShards));
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() {
      SimScratch sc;
      for (;;) { int sh = next.fetch_add(1); if (sh >= kShards) break;
        for (auto& kv : s->index[sh]) part[sh].push_back({similarity(s, c, sh, kv.second, sc), kv.first}); }
    });
    for (auto& t : th) t.join();
    for (auto& p : part) all.insert(all.end(), p.begin(), p.end());
  }
  // Total order T: ascending score, NaN after every number, then ascending id.
  // NEAREST = the first K of T; COLTT_COMPAT = the last K of T (the K largest distances,
  // SURVEY F1; among equal scores the higher ids stay, the exact mirror image); output in T order.
  auto isnan_ = [](float f) { return f != f; };
  auto t_less = [&](const HeapItem& a, const HeapItem& b) {
    bool an = isnan_(a.priority), bn = isnan_(b.priority);
    if (an != bn) return bn;
    if (!an) { if (a.priority < b.priority) return true; if (b.priority < a.priority) return false; }
    return a.id < b.id;
  };
  std::sort(all.begin(), all.end(), t_less);
  size_t k = std::min((size_t)top_k, all.size());
  if (select_mode == 1) all.resize(k);
  else all.erase(all.begin(), all.end() - k);
  return emit(all, out_ids, out_scores);
}

// Row access for tests (the stored, i.e. normalized+lowered, representation).
ORC_API int orc_store_get_row(OrcStore* s, uint64_t id, void* out) {
  int sh = (int)orc_shard_vertex(id, kShards);
  auto it = s->index[sh].find(id);
  if (it == s->index[sh].end()) return -1;
  size_t off = it->second * (size_t)s->dim;
  if (s->quant == Q_NONE) std::memcpy(out, &s->f32[sh][off], 4 * (size_t)s->dim);
  else if (s->quant == Q_F8 || s->quant == Q_F8E) std::memcpy(out, &s->f8[sh][off], s->dim);
  else std::memcpy(out, &s->f16[sh][off], 2 * (size_t)s->dim);
  return 0;
}

// ---------------------------------------------------------------------------
// SaveVertex / LoadVertex blob — none_vectorstore.go:308-516; element widths for the
// quantized stores per f16_vectorstore.go:338-343 / f8_vectorstore.go:340 (u16 / u8 BE).
// Vertices are written shard by shard in ascending id (reference: map order), metaCount 0.
// ---------------------------------------------------------------------------
static void put_be(std::vector<uint8_t>& b, uint64_t v, int nbytes) { for (int i = nbytes - 1; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i))); }

ORC_API size_t orc_store_save_vertex(OrcStore* s, uint8_t* out, size_t cap) {
  std::vector<uint8_t> b;
  for (int sh = 0; sh < kShards; sh++) {
    put_be(b, s->index[sh].size(), 8);
    for (auto& kv : s->index[sh]) {
      put_be(b, kv.first, 8);
      put_be(b, s->dim, 4);
      size_t off = kv.second * (size_t)s->dim;
      for (uint32_t i = 0; i < s->dim; i++) {
        if (s->quant == Q_NONE) put_be(b, f32bits(s->f32[sh][off + i]), 4);
        else if (s->quant == Q_F8) put_be(b, s->f8[sh][off + i], 1);
        else put_be(b, s->f16[sh][off + i], 2);
      }
      put_be(b, 0, 4);  // metaCount
    }
  }
  if (out && cap >= b.size()) std::memcpy(out, b.data(), b.size());
  return b.size();
}

// ---------------------------------------------------------------------------
// edge/resultset.go:42-119 — ResultSet (dead code in the reference, named by north_star)
// ---------------------------------------------------------------------------
struct OrcResultSet { std::vector<float> sims; std::vector<uint64_t> ids; int k; int valid; };
ORC_API OrcResultSet* orc_resultset_create(int k) { auto* r = new OrcResultSet(); r->k = k; r->sims.assign(k, 0.f); r->ids.assign(k, 0); r->valid = 0; return r; }
ORC_API void orc_resultset_destroy(OrcResultSet* r) { delete r; }
// resultset.go:71-108 AddResult
ORC_API int orc_resultset_add(OrcResultSet* rs, uint64_t id, float sim) {
  if (rs->valid == rs->k) { float last = rs->sims[rs->sims.size() - 1]; if (last > sim) return 0; }
  int insert = 0; bool found = false;
  while (insert != rs->k) {
    if (rs->valid <= insert) { rs->valid += 1; found = true; break; }
    if (rs->ids[insert] == id) return 1;
    if (rs->sims[insert] < sim) { found = true; break; }
    insert++;
  }
  if (!found) return 0;
  for (int i = rs->k - 1; i > insert; i--) { rs->sims[i] = rs->sims[i - 1]; rs->ids[i] = rs->ids[i - 1]; }  // copy(sims[insert+1:], sims[insert:])
  rs->sims[insert] = sim; rs->ids[insert] = id;
  return 1;
}
ORC_API int orc_resultset_to_slice(OrcResultSet* rs, uint64_t* ids, float* sims) {
  for (int i = 0; i < rs->valid; i++) { ids[i] = rs->ids[i]; sims[i] = rs->sims[i]; }
  return rs->valid;
}
// resultset.go:55-65 ComputeRecall
ORC_API double orc_compute_recall(const uint64_t* base_ids, const uint64_t* ids, int at) {
  int found = 0;
  for (int i = 0; i < at; i++) for (int j = 0; j < at; j++) if (base_ids[i] == ids[j]) found++;
  return (double)found / (double)at;
}

// ---------------------------------------------------------------------------
// core/vectorindex HNSW restatement
// ---------------------------------------------------------------------------
struct HVertex {
  uint64_t id; std::vector<float> vec; int level; bool deleted = false;
  std::vector<std::map<uint64_t, float>> edges;  // per level: neighbour id -> distance (hnsw_vertex.go:30)
};

struct OrcHnsw {
  uint32_t dim; int metric;
  // hnsw_config.go:135-162 defaults
  int search_algo = 0; float level_mult; int ef = 20, ef_construction = 200, m = 16, m_max, m_max0;
  bool heuristic_extend = false, heuristic_keep_pruned = true;
  std::map<uint64_t, std::unique_ptr<HVertex>> vertices;  // id order; shard = ShardVertex(id,16) at commit
  HVertex* entrypoint = nullptr;
  uint64_t stat_dist_evals = 0, stat_expansions = 0;
  SimScratch sc;
};

static float hdist(OrcHnsw* h, const float* a, const float* b) {
  h->stat_dist_evals++;
  size_t d = h->dim;
  if (g_ref_cos) {
    float* x = h->sc.bx.get(d); float* y = h->sc.by.get(d);
    std::memcpy(x, a, 4 * d); std::memcpy(y, b, 4 * d);
    return h->metric == M_COSINE ? cosine_distance_impl(d, x, y, true) : euclid_distance_impl(d, x, y, true);
  }
  return h->metric == M_COSINE ? cosine_distance_impl(d, a, b, false) : euclid_distance_impl(d, a, b, false);
}

ORC_API OrcHnsw* orc_hnsw_create(uint32_t dim, int metric, int m, int ef, int ef_construction, int heuristic) {
  OrcHnsw* h = new OrcHnsw();
  h->dim = dim; h->metric = metric;
  if (m > 0) h->m = m;
  if (ef > 0) h->ef = ef;
  if (ef_construction > 0) h->ef_construction = ef_construction;
  h->search_algo = heuristic ? 1 : 0;
  // levelMultiplier = 1/ln(m) through gomath.Log (float64 log -> f32), f32 divide
  h->level_mult = 1.0f / (float)std::log((double)(float)h->m);
  h->m_max = h->m; h->m_max0 = 2 * h->m;
  return h;
}
ORC_API void orc_hnsw_destroy(OrcHnsw* h) { delete h; }
ORC_API uint64_t orc_hnsw_len(OrcHnsw* h) { uint64_t n = 0; for (auto& kv : h->vertices) if (!kv.second->deleted) n++; return n; }
ORC_API void orc_hnsw_set_ef(OrcHnsw* h, int ef) { h->ef = ef; }
ORC_API void orc_hnsw_stats(OrcHnsw* h, uint64_t* evals, uint64_t* expansions, int reset) {
  *evals = h->stat_dist_evals; *expansions = h->stat_expansions;
  if (reset) { h->stat_dist_evals = 0; h->stat_expansions = 0; }
}
// hnsw.go:280-282 RandomLevel given U in (0,1): floor(-ln(U)*mL) via gomath (f64 log -> f32).
ORC_API int orc_hnsw_level_from_uniform(OrcHnsw* h, float u) {
  float l = -(float)std::log((double)u) * h->level_mult;
  return (int)std::floor((double)l);
}

static HVertex* hget(OrcHnsw* h, uint64_t id) { auto it = h->vertices.find(id); return it == h->vertices.end() ? nullptr : it->second.get(); }

// hnsw.go:320-343 greedyClosestNeighbor (neighbour iteration: ascending id)
static void greedy_closest(OrcHnsw* h, const float* q, HVertex*& ep, float& min_d, int level) {
  for (;;) {
    HVertex* closest = nullptr;
    for (auto& e : ep->edges[level]) {
      HVertex* nb = hget(h, e.first);
      if (!nb || nb->deleted) continue;
      float d = hdist(h, q, nb->vec.data());
      if (d < min_d) { min_d = d; closest = nb; }
    }
    if (!closest) break;
    ep = closest;
  }
}

// hnsw.go:345-389 searchLevel.  Returns the max-heap resultVertices.
static GoHeap search_level(OrcHnsw* h, const float* q, HVertex* ep, int ef, int level) {
  float epd = hdist(h, q, ep->vec.data());
  GoHeap cand(false), result(true);
  cand.push({epd, ep->id}); result.push({epd, ep->id});
  std::unordered_set<uint64_t> visited; visited.insert(ep->id);
  while (cand.len() > 0) {
    HeapItem ci = cand.pop();
    float lower_bound = result.peek().priority;
    if (ci.priority > lower_bound) break;
    HVertex* c = hget(h, ci.id);
    h->stat_expansions++;
    for (auto& e : c->edges[level]) {
      HVertex* nb = hget(h, e.first);
      if (!nb || nb->deleted) continue;
      if (!visited.insert(nb->id).second) continue;
      float d = hdist(h, q, nb->vec.data());
      if (d < lower_bound || (int)result.len() < ef) {
        cand.push({d, nb->id}); result.push({d, nb->id});
        if ((int)result.len() > ef) result.pop();
      }
    }
  }
  return result;
}

// hnsw.go:391-397 selectNeighbors
static void select_neighbors(GoHeap& nb, int k) { while ((int)nb.len() > k) nb.pop(); }

// hnsw.go:399-447 selectNeighborsHeuristic
static GoHeap select_neighbors_heuristic(OrcHnsw* h, const float* q, GoHeap& neighbors, int k, int level) {
  GoHeap cand(false);  // neighbors.Reverse(): min-queue over the same items (re-heapified)
  for (auto& it : neighbors.q) cand.push(it);
  std::unordered_set<uint64_t> existing;
  for (auto& it : neighbors.q) existing.insert(it.id);
  if (h->heuristic_extend) {
    while (neighbors.len() > 0) {
      HVertex* c = hget(h, neighbors.pop().id);
      for (auto& e : c->edges[level]) {
        HVertex* nb = hget(h, e.first);
        if (!nb || nb->deleted) continue;
        if (!existing.insert(nb->id).second) continue;
        cand.push({hdist(h, q, nb->vec.data()), nb->id});
      }
    }
  }
  GoHeap result(true);
  while (cand.len() > 0 && (int)result.len() < k) result.push(cand.pop());
  if (h->heuristic_keep_pruned)
    while (cand.len() > 0) { if ((int)result.len() >= k) break; result.push(cand.pop()); }
  return result;
}

// hnsw.go:449-474 pruneNeighbors
static void prune_neighbors(OrcHnsw* h, HVertex* v, int k, int level) {
  GoHeap nq(true);
  for (auto& e : v->edges[level]) { HVertex* nb = hget(h, e.first); if (!nb || nb->deleted) continue; nq.push({e.second, e.first}); }
  if (h->search_algo == 0) select_neighbors(nq, k);
  else nq = select_neighbors_heuristic(h, v->vec.data(), nq, k, level);
  std::map<uint64_t, float> ne;
  for (auto& it : nq.q) ne[it.id] = it.priority;
  v->edges[level] = std::move(ne);
}

// hnsw.go:104-167 Insert (single-threaded restatement; level supplied by the caller as in
// the reference's signature).
ORC_API int orc_hnsw_insert(OrcHnsw* h, uint64_t id, const float* value, int vertex_level) {
  if (h->vertices.count(id)) return -1;  // ItemAlreadyExistsError (hnsw.go:294-296)
  auto up = std::make_unique<HVertex>();
  HVertex* v = up.get();
  v->id = id; v->vec.resize(h->dim);
  if (h->metric == M_COSINE) orc_normalize(value, h->dim, v->vec.data());
  else std::memcpy(v->vec.data(), value, 4 * (size_t)h->dim);
  if (!h->entrypoint) {
    v->level = 0; v->edges.resize(1);
    h->vertices[id] = std::move(up);
    h->entrypoint = v;
    return 0;
  }
  v->level = vertex_level; v->edges.resize(vertex_level + 1);
  h->vertices[id] = std::move(up);
  HVertex* ep = h->entrypoint;
  float min_d = hdist(h, v->vec.data(), ep->vec.data());
  for (int l = ep->level; l > v->level; l--) greedy_closest(h, v->vec.data(), ep, min_d, l);
  for (int l = std::min(ep->level, v->level); l >= 0; l--) {
    GoHeap nbs = search_level(h, v->vec.data(), ep, h->ef_construction, l);
    if (h->search_algo == 0) select_neighbors(nbs, h->m);
    else nbs = select_neighbors_heuristic(h, v->vec.data(), nbs, h->m, l);
    int mmax = (l == 0) ? h->m_max0 : h->m_max;
    while (nbs.len() > 0) {
      HeapItem it = nbs.pop();
      HVertex* nb = hget(h, it.id);
      ep = nb;
      v->edges[l][nb->id] = it.priority;
      nb->edges[l][v->id] = it.priority;
      if ((int)nb->edges[l].size() > mmax) prune_neighbors(h, nb, mmax, l);
    }
  }
  if (h->entrypoint && v->level > h->entrypoint->level) h->entrypoint = v;
  return 0;
}

// hnsw.go:243-278 Search
ORC_API int orc_hnsw_search(OrcHnsw* h, const float* query, int k, uint64_t* out_ids, float* out_scores) {
  std::vector<float> qn(h->dim);
  if (h->metric == M_COSINE) orc_normalize(query, h->dim, qn.data());
  else std::memcpy(qn.data(), query, 4 * (size_t)h->dim);
  HVertex* ep = h->entrypoint;
  if (!ep) return 0;
  float min_d = hdist(h, qn.data(), ep->vec.data());
  for (int l = ep->level; l > 0; l--) greedy_closest(h, qn.data(), ep, min_d, l);
  int ef = std::max(h->ef, k);
  GoHeap nbs = search_level(h, qn.data(), ep, ef, 0);
  if (h->search_algo == 0) select_neighbors(nbs, k);
  else nbs = select_neighbors_heuristic(h, qn.data(), nbs, k, 0);
  int n = std::min(k, (int)nbs.len());
  for (int i = n - 1; i >= 0; i--) { HeapItem it = nbs.pop(); out_ids[i] = it.id; out_scores[i] = it.priority; }
  return n;
}

// hnsw_commit.go:69-162 Commit(header=true) with hnsw_config.go:179-201 save, metadata.go:31-41
// (empty metadata => uint16 0).  Shard = ShardVertex(id,16); vertices ascending id per shard.
ORC_API size_t orc_hnsw_commit(OrcHnsw* h, uint8_t* out, size_t cap) {
  std::vector<uint8_t> b;
  put_be(b, (uint32_t)h->search_algo, 4);
  put_be(b, f32bits(h->level_mult), 4);
  put_be(b, (uint32_t)h->ef, 4); put_be(b, (uint32_t)h->ef_construction, 4);
  put_be(b, (uint32_t)h->m, 4); put_be(b, (uint32_t)h->m_max, 4); put_be(b, (uint32_t)h->m_max0, 4);
  put_be(b, h->dim, 4);
  b.push_back(h->metric == M_COSINE ? 1 : 2);  // distToDistIdx
  std::vector<HVertex*> shards[kShards];
  for (auto& kv : h->vertices) if (!kv.second->deleted) shards[orc_shard_vertex(kv.first, kShards)].push_back(kv.second.get());
  size_t total = 0; for (auto& s : shards) total += s.size();
  if (total != 0 && h->entrypoint) {
    put_be(b, h->entrypoint->id, 8);
    for (auto& sh : shards) {
      put_be(b, sh.size(), 4);
      for (HVertex* v : sh) {
        put_be(b, v->id, 8); put_be(b, (uint32_t)v->level, 4);
        for (float f : v->vec) put_be(b, f32bits(f), 4);
        put_be(b, 0, 2);
      }
    }
    for (auto& sh : shards)
      for (HVertex* v : sh) {
        put_be(b, v->id, 8);
        for (int l = v->level; l >= 0; l--) {
          uint32_t cnt = 0;
          for (auto& e : v->edges[l]) { HVertex* nb = hget(h, e.first); if (nb && !nb->deleted) cnt++; }
          put_be(b, cnt, 4);
          for (auto& e : v->edges[l]) { HVertex* nb = hget(h, e.first); if (!nb || nb->deleted) continue; put_be(b, e.first, 8); put_be(b, f32bits(e.second), 4); }
        }
      }
  }
  if (out && cap >= b.size()) std::memcpy(out, b.data(), b.size());
  return b.size();
}

struct Reader {
  const uint8_t* p; size_t n, pos = 0; bool ok = true;
  uint64_t be(int nb) { if (pos + nb > n) { ok = false; return 0; } uint64_t v = 0; for (int i = 0; i < nb; i++) v = (v << 8) | p[pos++]; return v; }
  void skip(size_t k) { if (pos + k > n) ok = false; else pos += k; }
};

// hnsw_commit.go:164-278 Load(header=true)
ORC_API OrcHnsw* orc_hnsw_load(const uint8_t* data, size_t len) {
  Reader r{data, len};
  OrcHnsw* h = new OrcHnsw();
  h->search_algo = (int)r.be(4);
  h->level_mult = bitsf32((uint32_t)r.be(4));
  h->ef = (int32_t)r.be(4); h->ef_construction = (int32_t)r.be(4);
  h->m = (int32_t)r.be(4); h->m_max = (int32_t)r.be(4); h->m_max0 = (int32_t)r.be(4);
  h->dim = (uint32_t)r.be(4);
  uint8_t di = (uint8_t)r.be(1);
  if (!r.ok || (di != 1 && di != 2)) { delete h; return nullptr; }
  h->metric = di == 1 ? M_COSINE : M_EUCLID;
  if (r.pos == len) return h;  // empty index
  uint64_t ep_id = r.be(8);
  std::vector<std::vector<uint64_t>> shard_ids(kShards);
  for (int sh = 0; sh < kShards; sh++) {
    uint32_t cnt = (uint32_t)r.be(4);
    for (uint32_t i = 0; i < cnt && r.ok; i++) {
      auto v = std::make_unique<HVertex>();
      v->id = r.be(8); v->level = (int32_t)r.be(4);
      v->vec.resize(h->dim);
      for (uint32_t d = 0; d < h->dim; d++) v->vec[d] = bitsf32((uint32_t)r.be(4));
      uint16_t mc = (uint16_t)r.be(2);
      for (uint16_t k = 0; k < mc; k++) { size_t kl = r.be(1); r.skip(kl); size_t vl = r.be(2); r.skip(vl); }
      v->edges.resize(v->level + 1);
      shard_ids[sh].push_back(v->id);
      h->vertices[v->id] = std::move(v);
    }
  }
  h->entrypoint = hget(h, ep_id);
  for (int sh = 0; sh < kShards && r.ok; sh++)
    for (size_t i = 0; i < shard_ids[sh].size() && r.ok; i++) {
      uint64_t id = r.be(8);
      HVertex* v = hget(h, id);
      if (!v) { r.ok = false; break; }
      for (int l = v->level; l >= 0; l--) {
        uint32_t ne = (uint32_t)r.be(4);
        for (uint32_t j = 0; j < ne && r.ok; j++) { uint64_t nid = r.be(8); float d = bitsf32((uint32_t)r.be(4)); v->edges[l][nid] = d; }
      }
    }
  if (!r.ok) { delete h; return nullptr; }
  return h;
}

// hnsw.go:188-241 Remove (entrypoint hand-over + neighbour re-prune)
ORC_API int orc_hnsw_remove(OrcHnsw* h, uint64_t id) {
  HVertex* v = hget(h, id);
  if (!v || v->deleted) return -1;
  v->deleted = true;
  if (h->entrypoint == v) {
    float min_d = 3.402823466e+38f; HVertex* closest = nullptr;
    for (int l = v->level; l >= 0; l--) {
      for (auto& e : v->edges[l]) if (e.second < min_d) { min_d = e.second; closest = hget(h, e.first); }
      if (closest) break;
    }
    h->entrypoint = closest;
  }
  for (int l = v->level; l >= 0; l--) {
    int mmax = (l == 0) ? h->m_max0 : h->m_max;
    std::vector<uint64_t> nbs; for (auto& e : v->edges[l]) nbs.push_back(e.first);
    for (uint64_t nid : nbs) { HVertex* nb = hget(h, nid); if (!nb) continue; nb->edges[l].erase(v->id); prune_neighbors(h, nb, mmax, l); }
  }
  return 0;
}
