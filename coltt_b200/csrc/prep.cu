// prep.cu — ingest / query preparation: Normalize -> Lower -> AVX-order ||x||^2, on the GPU.
//
// Reference path being replaced (per vector, on the CPU):
//   edge.Normalize                    edge/vectorstore.go:173-189  (sequential f32 sum, sqrt via f64, f32 divide)
//   Quantization.Lower                edge/{f16,bf16,f8}_quantization.go  -> pkg/compresshelper
//   ||x||^2 in the cosine kernel      pkg/distance/simd/cpp/avx.cpp:51-75 (recomputed per distance call there;
//                                     here computed once per row at ingest, same lane order, 4 B/row)
// One warp per vector.  The Normalize sum is inherently sequential (its rounding sequence is
// part of the parity contract), so lane 0 walks the row out of shared memory; everything
// else (including the rounded squares the chain adds up) is lane-parallel.  HBM traffic: reads dim*4 B, writes dim*elem B (+ optional fp32/fp16
// query copies) per vector — an ingest-time cost, not on the search path.
#include "kernels.cuh"
#include "store.h"

namespace coltt {

template <int ELEM>
__global__ void __launch_bounds__(256) prep_rows_kernel(PrepParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dim = p.dim;
  float* x = smem_f + (size_t)warp * 2 * p.smem_stride;
  float* sq = x + p.smem_stride;   // x[d]*x[d], rounded: the products are lane-parallel, only the adds are a chain

  // Scratch initialisation for the search that follows (the FAST path's bound / count / published-maximum
  // arrays): folded into this launch so that no memset nodes sit between the query prep and the scan.
#pragma unroll
  for (int f = 0; f < 3; f++) {
    uint32_t* fp = p.fill_ptr[f];
    if (fp)
      for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < p.fill_words[f]; w += (size_t)gridDim.x * blockDim.x) fp[w] = p.fill_value[f];
  }

  for (size_t i = (size_t)blockIdx.x * warps_per_block + warp; i < p.n; i += (size_t)gridDim.x * warps_per_block) {
    const float* src = p.in + i * (size_t)p.in_stride;
    for (uint32_t d = lane; d < dim; d += 32) { const float v = src[d]; x[d] = v; sq[d] = mul_rn(v, v); }
    for (uint32_t d = dim + lane; d < p.smem_stride; d += 32) sq[d] = 0.0f;
    __syncwarp();

    if (p.normalize) {
      float norm = 0.0f;
      if (lane == 0) {
        // edge/vectorstore.go:176-178: for i := range v { norm += v[i] * v[i] }  (unfused, in order).
        // The rounded products were formed above by all lanes; this chain is the sequence of adds only.
        const uint32_t d4 = dim / 4 * 4;
        for (uint32_t d = 0; d < d4; d += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sq + d);
          norm = add_rn(add_rn(add_rn(add_rn(norm, s4.x), s4.y), s4.z), s4.w);
        }
        for (uint32_t d = d4; d < dim; d++) norm = add_rn(norm, sq[d]);
      }
      norm = __shfl_sync(0xffffffffu, norm, 0);
      if (norm == 0.0f) {
        for (uint32_t d = lane; d < dim; d += 32) x[d] = 0.0f;  // :179-181 zero vector stays zero
      } else {
        float nrm = sqrt_via_f64(norm);                            // :183
        for (uint32_t d = lane; d < dim; d += 32) x[d] = __fdiv_rn(x[d], nrm);  // :184-186
      }
      __syncwarp();
    }

    // Lower + write the stored row; keep the dequantized value in smem for the norm pass.
    const size_t slot = p.slots ? (size_t)p.slots[i] : (size_t)p.slot_base + i;
    uint8_t* row = p.rows_out ? p.rows_out + slot * (size_t)p.row_stride : nullptr;
    float e_scale = 1.0f, e_inv = 1.0f;
    if (ELEM == ELEM_F8E) {
      // builder-defined E4M3 store: one power-of-two scale per vector from max|v| (common.cuh, include/coltt_b200.h)
      float mx = 0.0f;
      bool bad = false;
      for (uint32_t d = lane; d < dim; d += 32) {
        const float a = fabsf(x[d]);
        if (!(a <= 3.402823466e+38f)) bad = true; else mx = fmaxf(mx, a);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      bad = __any_sync(0xffffffffu, bad);
      e_scale = bad ? 1.0f : e4m3_scale_from_maxabs(mx);
      e_inv = __uint_as_float((uint32_t)(254 - (int)(__float_as_uint(e_scale) >> 23)) << 23);   // 1 / 2^es, exact
      if (lane == 0 && p.scale_out) p.scale_out[p.norm2_by_slot ? slot : i] = e_scale;
    }
    uint8_t* code_row = (ELEM == ELEM_F8E && p.code_out) ? p.code_out + i * (size_t)p.code_stride : nullptr;
    for (uint32_t d = lane; d < dim; d += 32) {
      float v = x[d];
      if (ELEM == ELEM_F8E) {
        const uint8_t c = e4m3_encode(mul_rn(v, e_inv));
        if (row) row[d] = c;
        if (code_row) code_row[d] = c;
        v = mul_rn(e_scale, e4m3_decode(c));
      } else if (ELEM == ELEM_F32) {
        if (row) reinterpret_cast<float*>(row)[d] = v;
        if (p.shadow_out) reinterpret_cast<__half*>(p.shadow_out + slot * (size_t)p.shadow_stride)[d] = __float2half_rn(v);
      } else if (ELEM == ELEM_F16) {
        // compresshelper.Fromfloat32 (float16.go:124,274-321): IEEE RNE incl. subnormals and
        // overflow->inf == __float2half_rn for every non-NaN input (tests/test_gpu_codec.py).
        __half h = __float2half_rn(v);
        if (row) reinterpret_cast<__half*>(row)[d] = h;
        v = __half2float(h);  // Float16.Float32 (float16.go:184,237-270), exact
      } else {
        uint8_t c = f8_compat_encode(__float_as_uint(v));   // F8Fromfloat32 (float8.go:120,270-313)
        if (row) row[d] = c;
        v = __uint_as_float(f8_compat_decode_bits(c));       // Float8.Float32 (float8.go:180,233-266)
      }
      x[d] = v;
    }
    // zero the padding of the stored row so bulk copies of whole 16-byte units are defined
    if (row) {
      const uint32_t es = ELEM == ELEM_F32 ? 4 : (ELEM == ELEM_F16 ? 2 : 1);
      for (uint32_t b = dim * es + lane; b < p.row_stride; b += 32) row[b] = 0;
    }
    if (code_row)
      for (uint32_t b = dim + lane; b < p.code_stride; b += 32) code_row[b] = 0;   // +0.0 in E4M3
    if (ELEM == ELEM_F32 && p.shadow_out)
      for (uint32_t b = dim * 2 + lane; b < p.shadow_stride; b += 32) p.shadow_out[slot * (size_t)p.shadow_stride + b] = 0;
    __syncwarp();

    // ||x||^2 exactly as cosine_similarity_dot_norm accumulates it for one operand
    // (avx.cpp:57-63 lanes, :3-8 tree, :68-72 scalar tail).
    float acc = 0.0f;
    const uint32_t full = (dim / 8) * 8;
    if (lane < 8)
      for (uint32_t d = lane; d < full; d += 8) acc = add_rn(acc, mul_rn(x[d], x[d]));
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));  // (l0+l1) (l2+l3) (l4+l5) (l6+l7)
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));  // (l0+l1)+(l2+l3), (l4+l5)+(l6+l7)
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));  // sum of both halves
    if (lane == 0) {
      for (uint32_t d = full; d < dim; d++) acc = add_rn(acc, mul_rn(x[d], x[d]));
      if (p.norm2_out) p.norm2_out[p.norm2_by_slot ? slot : i] = acc;
    }
    if (p.deq_out) {
      float* dq = p.deq_out + i * (size_t)p.deq_stride;
      for (uint32_t d = lane; d < p.deq_stride; d += 32) dq[d] = d < dim ? x[d] : 0.0f;
    }
    if (p.f16_out) {
      __half* hq = p.f16_out + i * (size_t)p.f16_stride;
      for (uint32_t d = lane; d < p.f16_stride; d += 32) hq[d] = d < dim ? __float2half_rn(x[d]) : __half(0);
    }
    __syncwarp();
  }
}

int launch_prep_rows(const PrepParams& p, int elem, cudaStream_t stream) {
  if (p.n == 0) return COLTT_OK;
  // warps per block bounded by shared memory: one fp32 copy of the vector per warp
  if (p.smem_stride % 4) return fail(COLTT_ERR_INVALID, "prep: smem_stride must be a multiple of 4 floats");
  const size_t per_warp = (size_t)p.smem_stride * sizeof(float) * 2;   // the vector and its squares
  int warps = 8;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  if (per_warp * warps > 227 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "dim too large for the ingest kernel");
  const size_t smem = per_warp * warps;
  const size_t blocks_needed = (p.n + warps - 1) / warps;
  const int grid = (int)(blocks_needed < 148 * 8 ? blocks_needed : 148 * 8);
  auto k = elem == ELEM_F32 ? prep_rows_kernel<ELEM_F32>
                            : (elem == ELEM_F16 ? prep_rows_kernel<ELEM_F16> : (elem == ELEM_F8C ? prep_rows_kernel<ELEM_F8C> : prep_rows_kernel<ELEM_F8E>));
  { int arc = kernel_attrs(k, smem); if (arc) return arc; }
  k<<<grid, warps * 32, smem, stream>>>(p);
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
