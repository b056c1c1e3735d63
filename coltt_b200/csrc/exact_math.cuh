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
  } else {
    uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
    v[0] = lut[raw & 0xff]; v[1] = lut[(raw >> 8) & 0xff]; v[2] = lut[(raw >> 16) & 0xff]; v[3] = lut[raw >> 24];
  }
}
template <int ELEM>
__device__ __forceinline__ float load1(const uint8_t* row, uint32_t idx, const float* lut) {
  if (ELEM == ELEM_F32) return reinterpret_cast<const float*>(row)[idx];
  if (ELEM == ELEM_F16) return __half2float(reinterpret_cast<const __half*>(row)[idx]);
  return lut[row[idx]];
}


// lane -> (row within a 16-row group, half g): the 8 lanes of each quarter-warp touch 8 different rows
__device__ __forceinline__ uint32_t lane_row16(uint32_t lane) { return (lane & 7) + 8 * (lane >> 4); }
__device__ __forceinline__ uint32_t lane_half(uint32_t lane) { return (lane >> 3) & 1; }

}  // namespace coltt
