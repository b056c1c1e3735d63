// blob.cu — the reference's on-disk vertex blob as a GPU ingest/egress format.
//   SaveVertex  edge/none_vectorstore.go:308-423  (element widths: f16_vectorstore.go:338-343 u16,
//               f8_vectorstore.go:340 u8, all big-endian)
//   LoadVertex  edge/none_vectorstore.go:425-516
// Layout: for shard 0..15 { u64 count; count x { u64 id; u32 vecLen; vecLen elements;
//         u32 metaCount; metaCount x { u16 keyLen; key; u8 tag; value } } }.
// Rows hold the STORED representation (normalized + lowered), so import bypasses
// Normalize/Lower and only recomputes ||row||^2 on the device.
#include <algorithm>
#include <cstring>

#include "kernels.cuh"
#include "hnsw.h"
#include "store.h"

namespace coltt {

template <int ELEM>
__global__ void __launch_bounds__(256) norm2_stored_kernel(const uint8_t* rows, uint32_t row_stride, uint32_t dim, size_t n,
                                                          float* norm2) {
  // same lane order as prep.cu / avx.cpp:57-63; one warp per stored row, lanes 0..7 = AVX lanes
  const uint32_t lane = threadIdx.x & 31;
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const uint8_t* row = rows + w * row_stride;
  auto val = [&](uint32_t d) -> float {
    if (ELEM == ELEM_F32) return reinterpret_cast<const float*>(row)[d];
    if (ELEM == ELEM_F16) return __half2float(reinterpret_cast<const __half*>(row)[d]);
    return __uint_as_float(f8_compat_decode_bits(row[d]));
  };
  const uint32_t full = (dim / 8) * 8;
  float acc = 0.0f;
  if (lane < 8)
    for (uint32_t d = lane; d < full; d += 8) { float v = val(d); acc = add_rn(acc, mul_rn(v, v)); }
  acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
  acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
  acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
  if (lane == 0) {
    for (uint32_t d = full; d < dim; d++) { float v = val(d); acc = add_rn(acc, mul_rn(v, v)); }
    norm2[w] = acc;
  }
}

__global__ void shadow_from_rows_kernel(const uint8_t* rows, uint32_t row_stride, uint32_t dim, size_t n, uint8_t* shadow, uint32_t shadow_stride) {
  // fp16 shadow of stored fp32 rows (store.cu): element-wise round to nearest, padding zeroed
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= n) return;
  const float* src = reinterpret_cast<const float*>(rows + w * row_stride);
  __half* dst = reinterpret_cast<__half*>(shadow + w * (size_t)shadow_stride);
  for (uint32_t d = lane; d < shadow_stride / 2; d += 32) dst[d] = d < dim ? __float2half_rn(src[d]) : __half(0);
}

int launch_norm2_stored_f32(const uint8_t* rows, uint32_t row_stride, uint32_t dim, size_t n, float* norm2, cudaStream_t stream) {
  if (n == 0) return COLTT_OK;
  const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
  norm2_stored_kernel<ELEM_F32><<<blocks, 256, 0, stream>>>(rows, row_stride, dim, n, norm2);
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

static inline void put_be(uint8_t*& p, uint64_t v, int nb) {
  for (int i = nb - 1; i >= 0; i--) *p++ = (uint8_t)(v >> (8 * i));
}

int Store::export_blob(void* buf, size_t* len) {
  std::shared_lock<std::shared_mutex> lk(mu);
  if (elem == ELEM_F8E) return fail(COLTT_ERR_UNSUPPORTED, "SaveVertex has no layout for the F8_E4M3 extension (per-row scales)");
  if (anonymous) return fail(COLTT_ERR_UNSUPPORTED, "store was filled from device memory: no host id map to export");
  const uint32_t es = elem_size(elem);
  const size_t per = 8 + 4 + (size_t)dim * es + 4;
  const size_t need = 16 * 8 + n_rows * per;
  if (!buf) { *len = need; return COLTT_OK; }
  if (*len < need) { *len = need; return fail(COLTT_ERR_INVALID, "export buffer too small"); }
  COLTT_CUDA(cudaSetDevice(device));
  std::vector<uint8_t> host(n_rows * (size_t)row_stride);
  if (n_rows) COLTT_CUDA(cudaMemcpy(host.data(), d_rows, host.size(), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> order[16];
  for (size_t s = 0; s < n_rows; s++) order[shard_vertex(h_ids[s], 16)].push_back((uint32_t)s);
  uint8_t* p = (uint8_t*)buf;
  for (int sh = 0; sh < 16; sh++) {
    std::sort(order[sh].begin(), order[sh].end(), [&](uint32_t a, uint32_t b) { return h_ids[a] < h_ids[b]; });
    put_be(p, order[sh].size(), 8);
    for (uint32_t s : order[sh]) {
      put_be(p, h_ids[s], 8);
      put_be(p, dim, 4);
      const uint8_t* row = host.data() + (size_t)s * row_stride;
      for (uint32_t d = 0; d < dim; d++) {  // little-endian device element -> big-endian blob element
        for (int b = (int)es - 1; b >= 0; b--) *p++ = row[(size_t)d * es + b];
      }
      put_be(p, 0, 4);  // metaCount: metadata stays on the Go side
    }
  }
  *len = (size_t)(p - (uint8_t*)buf);
  return COLTT_OK;
}

struct BlobReader {
  const uint8_t* p; size_t n, pos = 0; bool ok = true;
  uint64_t be(int nb) {
    if (pos + nb > n) { ok = false; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v = (v << 8) | p[pos++];
    return v;
  }
  void skip(size_t k) { if (pos + k > n) ok = false; else pos += k; }
};

int Store::import_blob(const void* buf, size_t len) {
  std::unique_lock<std::shared_mutex> lk(mu);
  if (elem == ELEM_F8E) return fail(COLTT_ERR_UNSUPPORTED, "LoadVertex has no layout for the F8_E4M3 extension (per-row scales)");
  COLTT_CUDA(cudaSetDevice(device));
  { int wrc = wait_for_searches(); if (wrc) return wrc; }
  const uint32_t es = elem_size(elem);
  BlobReader r{(const uint8_t*)buf, len};
  std::vector<uint64_t> ids;
  std::vector<uint8_t> rows;
  for (int sh = 0; sh < 16 && r.ok; sh++) {
    const uint64_t count = r.be(8);
    for (uint64_t j = 0; j < count && r.ok; j++) {
      const uint64_t id = r.be(8);
      const uint32_t vlen = (uint32_t)r.be(4);
      if (!r.ok) break;
      if (vlen != dim) return fail(COLTT_ERR_DIM, "vertex blob dimension " + std::to_string(vlen) + " != collection dim " + std::to_string(dim));
      if (r.pos + (size_t)vlen * es > r.n) { r.ok = false; break; }
      const size_t off = rows.size();
      rows.resize(off + row_stride, 0);
      const uint8_t* src = r.p + r.pos;
      for (uint32_t d = 0; d < dim; d++)
        for (uint32_t b = 0; b < es; b++) rows[off + (size_t)d * es + b] = src[(size_t)d * es + (es - 1 - b)];
      r.skip((size_t)vlen * es);
      const uint32_t mc = (uint32_t)r.be(4);
      for (uint32_t m = 0; m < mc && r.ok; m++) {
        r.skip(r.be(2));
        const uint8_t tag = (uint8_t)r.be(1);
        if (tag == 0 || tag == 2) r.skip(8);
        else if (tag == 1) r.skip(r.be(2));
        else if (tag == 3) r.skip(1);
        else return fail(COLTT_ERR_FORMAT, "unsupported metadata type tag: " + std::to_string(tag));  // none_vectorstore.go:503
      }
      ids.push_back(id);
    }
  }
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated vertex blob");
  // LoadVertex replaces the whole collection (none_vectorstore.go:509-513)
  n_rows = 0; h_ids.clear(); id2slot.clear();
  std::unordered_map<uint64_t, uint32_t> seen;
  std::vector<uint32_t> keep;
  for (size_t i = 0; i < ids.size(); i++) {  // a map load keeps the last occurrence of a repeated key
    auto it = seen.find(ids[i]);
    if (it != seen.end()) keep[it->second] = (uint32_t)i;
    else { seen[ids[i]] = (uint32_t)keep.size(); keep.push_back((uint32_t)i); }
  }
  const size_t n = keep.size();
  int rc = reserve(n);
  if (rc) return rc;
  if (n) {
    if (n != ids.size()) {
      std::vector<uint8_t> packed(n * (size_t)row_stride);
      for (size_t j = 0; j < n; j++) std::memcpy(&packed[j * row_stride], &rows[(size_t)keep[j] * row_stride], row_stride);
      rows.swap(packed);
    }
    h_ids.resize(n);
    for (size_t j = 0; j < n; j++) { h_ids[j] = ids[keep[j]]; id2slot[h_ids[j]] = (uint32_t)j; }
    COLTT_CUDA(cudaMemcpyAsync(d_rows, rows.data(), n * (size_t)row_stride, cudaMemcpyHostToDevice, stream));
    COLTT_CUDA(cudaMemcpyAsync(d_ids, h_ids.data(), n * 8, cudaMemcpyHostToDevice, stream));
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    if (elem == ELEM_F32) norm2_stored_kernel<ELEM_F32><<<blocks, 256, 0, stream>>>(d_rows, row_stride, dim, n, d_norm2);
    else if (elem == ELEM_F16) norm2_stored_kernel<ELEM_F16><<<blocks, 256, 0, stream>>>(d_rows, row_stride, dim, n, d_norm2);
    else norm2_stored_kernel<ELEM_F8C><<<blocks, 256, 0, stream>>>(d_rows, row_stride, dim, n, d_norm2);
    if (d_shadow) shadow_from_rows_kernel<<<blocks, 256, 0, stream>>>(d_rows, row_stride, dim, n, d_shadow, shadow_stride);
    COLTT_CUDA(cudaGetLastError());
    COLTT_CUDA(cudaStreamSynchronize(stream));
  }
  n_rows = n;
  return COLTT_OK;
}

}  // namespace coltt
