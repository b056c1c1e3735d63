// common.cuh — shared device/host helpers for libcoltt_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>

#include "../../include/coltt_b200.h"

namespace coltt {

// ---- error plumbing (never throws across the C-ABI) -----------------------------------
void set_last_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define COLTT_CUDA(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::coltt::fail(COLTT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ---- result record exchanged between kernels and across GPUs (== coltt_hit) ------------
struct __align__(16) Hit {
  uint64_t id;
  float score;
  uint32_t slot;
};
static_assert(sizeof(Hit) == 16 && sizeof(coltt_hit) == 16, "Hit layout");

// ---- total order T on (score, id): ascending score, NaN after every number, then id ----
// NEAREST keeps the first K of T, COLTT_COMPAT keeps the last K of T (edge/priority_queue.go
// keeps the K largest); both are reported in T order.
__host__ __device__ __forceinline__ bool t_less(float sa, uint64_t ia, float sb, uint64_t ib) {
  bool an = sa != sa, bn = sb != sb;
  if (an | bn) {
    if (an != bn) return bn;
    return ia < ib;
  }
  if (sa != sb) return sa < sb;
  return ia < ib;
}
// "a ranks better than b" for the given select mode.
__host__ __device__ __forceinline__ bool better(float sa, uint64_t ia, float sb, uint64_t ib, int nearest) {
  return nearest ? t_less(sa, ia, sb, ib) : t_less(sb, ib, sa, ia);
}

// ---- exact (reference-order) arithmetic -------------------------------------------------
// The Go/AVX path never fuses (no FMA in pkg/distance/simd/avx/AVX_amd64.s; GOAMD64=v1 for
// the Go loops), so every op here is an explicit round-to-nearest intrinsic that nvcc may not
// contract.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
// float32(math.Sqrt(float64(x)))  (simd/avx/AVX_amd64.go:31,51; edge/vectorstore.go:184)
__device__ __forceinline__ float sqrt_via_f64(float x) { return __double2float_rn(__dsqrt_rn((double)x)); }

// simd/avx/AVX_amd64.go:44-52 CosineDistance + space.go:93-95 (gomath.Abs)
__device__ __forceinline__ float cosine_epilogue(float dot, float na, float nb) {
  float n2 = mul_rn(na, nb);                    // avx.cpp:74  norm_a_sum * norm_b_sum
  float q = __fdiv_rn(dot, sqrt_via_f64(n2));   // dot / float32(sqrt(float64(n2)))
  return fabsf(sub_rn(1.0f, q));
}

// experimental/experimental_helper.go:134-139 scoreHelper (== edge/edge_helper.go:143-148): cosine ((2 - s) / 2) * 100 in
// float32; euclidean float32(math.Max(0, float64(100 - s))) — NaN stays NaN, -0 becomes +0.
__device__ __forceinline__ float score_helper(float s, int metric) {
  if (metric == COLTT_COSINE) return mul_rn(__fdiv_rn(sub_rn(2.0f, s), 2.0f), 100.0f);
  const float x = sub_rn(100.0f, s);
  return x != x ? x : (x > 0.0f ? x : 0.0f);
}

// ---- codecs ----------------------------------------------------------------------------
// pkg/compresshelper/float8.go:270-313 f32bitsToF8bits, literal.
__host__ __device__ __forceinline__ uint8_t f8_compat_encode(uint32_t u32) {
  uint32_t sign = u32 & 0x800000u;
  uint32_t exp = u32 & 0x7f800000u;
  uint32_t coef = u32 & 0x007fffffu;
  if (exp == 0x7f800000u) {
    uint32_t nanBit = coef != 0 ? 0x0200u : 0u;
    return (uint8_t)((sign >> 8) | 0x7cu | nanBit | (coef >> 13));
  }
  uint32_t halfSign = sign >> 8;
  int32_t halfExp = (int32_t)(exp >> 23) - 127 + 15;
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
  if ((coef & roundBit) != 0 && (coef & (3 * roundBit - 1)) != 0) return (uint8_t)((halfSign | uHalfExp | halfCoef) + 1);
  return (uint8_t)(halfSign | uHalfExp | halfCoef);
}
// pkg/compresshelper/float8.go:233-266 F8bitsToF32bits, literal: exp field is always 0.
__host__ __device__ __forceinline__ uint32_t f8_compat_decode_bits(uint8_t in) {
  uint32_t sign = (uint32_t)(in & 0x80) << 8;
  uint32_t coef = (uint32_t)(in & 0x03) << 13;
  if (coef == 0) return sign;
  uint32_t exp = 1;
  while ((coef & 0x7f800000u) == 0) {
    coef <<= 1;
    exp--;
  }
  coef &= 0x007fffffu;
  return sign | ((exp + (0x7f - 0xf)) << 23) | coef;
}

// F8_E4M3 store (builder-defined extension, no reference arithmetic — SURVEY F3; include/coltt_b200.h states the
// format): OCP E4M3 "fn" codes, RNE, saturating, with one power-of-two scale per vector:
//   s = 2^clamp(floor(log2(max|v|)) - 7, -40, 40)  (1 when max|v| is 0 or not finite),  c_i = e4m3(v_i / s),  x_i = s * dec(c_i).
__device__ __forceinline__ float e4m3_scale_from_maxabs(float mx) {
  if (!(mx <= 3.402823466e+38f) || mx == 0.0f) return 1.0f;
  int e = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127;          // floor(log2(mx)) for normal mx
  if (((__float_as_uint(mx) >> 23) & 0xff) == 0) e = -127;          // subnormal: clamps below anyway
  int es = e - 7;
  es = es < -40 ? -40 : (es > 40 ? 40 : es);
  return __uint_as_float((uint32_t)(es + 127) << 23);
}
__device__ __forceinline__ uint8_t e4m3_encode(float x) {   // cvt.rn.satfinite.e4m3x2.f32
  return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float2 e4m3x2_decode(uint16_t c2) {   // cvt.rn.f16x2.e4m3x2: exact
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)c2, __NV_E4M3);
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
__device__ __forceinline__ float e4m3_decode(uint8_t c) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)c, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}

// ---- order-preserving float <-> uint32 (atomicMin/Max on bounds that may be negative) -------------
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

// ---- mbarrier / bulk-copy (TMA) PTX ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe: the answer arrives like a load result, so it can be requested before other work and consumed after it
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace coltt
