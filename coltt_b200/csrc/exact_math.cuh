// exact_math.cuh — element decode shared by the exact-order kernels (flat_scan.cu, rerank.cu).
// A thread owns AVX lanes 4g..4g+3 of a row (pkg/distance/simd/cpp/avx.cpp:15-32,51-75): it reads
// 4 consecutive elements out of every group of 8 and keeps 4 independent accumulation chains.
#pragma once
#include "kernels.cuh"

namespace coltt {

template <int ELEM>
__device__ __forceinline__ void load4(const uint8_t* p, const float* lut, float (&v)[4]) {
  if (ELEM == ELEM_F32) {
    float4 f = *reinterpret_cast<const float4*>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  } else if (ELEM == ELEM_F16) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else if (ELEM == ELEM_F8E) {
    // E4M3 codes -> the UNSCALED decoded values (cvt.rn.f16x2.e4m3x2, exact); the caller applies the row's
    // power-of-two scale (per element for L2, once per row for the dot product — same bits, see flat_scan.cu)
    uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
    const float2 a = e4m3x2_decode((uint16_t)(raw & 0xffffu)), b = e4m3x2_decode((uint16_t)(raw >> 16));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    // reference f8 codes (float8.go:233-266): the decoder reads only bits 0-1 and bit 7 of a code, so (code & 0x83)
    // indexes a table that every lane keeps a PRIVATE copy of, one bank per lane (lut = table + lane, entries 32 floats
    // apart): the four lookups are bank-conflict free whatever the codes are.  (A shared 256-entry table cost ~3.5
    // conflicting wavefronts per lookup on random codes: 0.22 of the HBM roofline.)
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(p) & 0x83838383u;
    v[0] = lut[(raw & 0xffu) * 32]; v[1] = lut[((raw >> 8) & 0xffu) * 32]; v[2] = lut[((raw >> 16) & 0xffu) * 32]; v[3] = lut[(raw >> 24) * 32];
  }
}
template <int ELEM>
__device__ __forceinline__ float load1(const uint8_t* row, uint32_t idx, const float* lut) {
  if (ELEM == ELEM_F32) return reinterpret_cast<const float*>(row)[idx];
  if (ELEM == ELEM_F16) return __half2float(reinterpret_cast<const __half*>(row)[idx]);
  if (ELEM == ELEM_F8E) return e4m3_decode(row[idx]);
  return lut[(uint32_t)(row[idx] & 0x83u) * 32];
}


// dot += q*r exactly as the reference's unfused `dot = dot + v1*v2` (avx.cpp:60) rounds it.  On the fp16 ("bf16")
// stores both operands are decoded binary16 values: at most 11 significant bits each and magnitudes >= 2^-24,
// so q*r has <= 22 significant bits and lies far inside the fp32 normal range — the product is exact in fp32
// and round(q*r + acc) == round(round(q*r) + acc).  One FFMA therefore produces the reference's bits at half
// the issue cost (tests/test_oracle.py::test_fp16_products_are_exact_in_fp32 pins the premise).
// fp32 rows keep the two roundings, and so does the f8-compat store: its decoder leaves a stray mantissa bit and
// fp32-subnormal values for codes >= 0x80 (float8.go:233-266), whose products underflow.
// (L2's (q-r)^2 has no such property: q-r may need > 24 bits.)
// The E4M3 store has the same property with room to spare: both operands are (power of two) x (4 significant bits),
// scales are clamped to 2^+-40, so every product has <= 8 significant bits and is exact.
template <int ELEM>
__device__ __forceinline__ float dot_step(float acc, float q, float r) {
  if (ELEM == ELEM_F16 || ELEM == ELEM_F8E) return __fmaf_rn(q, r, acc);
  return add_rn(acc, mul_rn(q, r));
}

// lane -> (row within a 16-row group, half g): the 8 lanes of each quarter-warp touch 8 different rows
__device__ __forceinline__ uint32_t lane_row16(uint32_t lane) { return (lane & 7) + 8 * (lane >> 4); }
__device__ __forceinline__ uint32_t lane_half(uint32_t lane) { return (lane >> 3) & 1; }

}  // namespace coltt
