// topk_merge.cu — K5: final merge of per-CTA (or per-GPU) top-K lists into the answer.
//
// Replaces the tail of VertexSearch: re-Adding the 16 shard-local queues into the global
// queue and PriorityQueue.ToSlice's sort (edge/none_vectorstore.go:173-179,
// edge/priority_queue.go:57-69); across GPUs it is the one exchange step of the sharded search
// (after the all-gather of per-shard lists).  One CTA per query; rank-based
// selection (topk.cuh) — no sort, output written directly in T order.
#include "kernels.cuh"
#include "store.h"
#include "topk.cuh"

namespace coltt {

template <bool STAGED>
__global__ void __launch_bounds__(256) merge_topk_kernel(MergeParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t q = blockIdx.x;
  const int rev = (!p.in_best_first && !p.nearest) ? 1 : 0;  // public T-order lists are worst-first for COMPAT
  Hit* sel = reinterpret_cast<Hit*>(smem);                       // [k]
  const Hit* L;
  const int* cnt;
  size_t list_stride, cnt_stride;
  if (STAGED) {
    Hit* Ls = sel + p.k;                                         // [n_lists][k_in]
    int* cnt_s = reinterpret_cast<int*>(Ls + (size_t)p.n_lists * p.k_in);
    for (int j = threadIdx.x; j < p.n_lists; j += blockDim.x) {
      int c = p.counts[(size_t)j * p.nq + q];
      cnt_s[j] = c > (int)p.k_in ? (int)p.k_in : c;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (uint32_t)p.n_lists * p.k_in; i += blockDim.x) {
      uint32_t j = i / p.k_in, e = i - j * p.k_in;
      if ((int)e < cnt_s[j]) Ls[i] = p.lists[((size_t)j * p.nq + q) * p.k_in + e];
    }
    __syncthreads();
    L = Ls; cnt = cnt_s; list_stride = p.k_in; cnt_stride = 1;
  } else {
    L = p.lists + (size_t)q * p.k_in; cnt = p.counts + q;
    list_stride = (size_t)p.nq * p.k_in; cnt_stride = p.nq;
  }
  rank_merge_block(L, cnt, p.n_lists, p.k_in, p.k, p.nearest, sel, rev, list_stride, cnt_stride);
  __syncthreads();
  uint32_t total = 0;
  for (int j = 0; j < p.n_lists; j++) {
    uint32_t c = (uint32_t)cnt[(size_t)j * cnt_stride];
    total += c > p.k_in ? p.k_in : c;
  }
  const uint32_t n_out = total < p.k ? total : p.k;
  Hit* out = p.out + (size_t)q * p.k;
  // sel is best-first; T order is the same for NEAREST and the exact mirror for COMPAT.
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) out[p.nearest ? i : n_out - 1 - i] = sel[i];
  if (threadIdx.x == 0) p.out_counts[q] = (int)n_out;
}

int launch_merge_topk(const MergeParams& p, cudaStream_t stream) {
  if (p.nq == 0) return COLTT_OK;
  if (p.k == 0 || p.k > 1024 || p.k_in == 0) return fail(COLTT_ERR_UNSUPPORTED, "merge: k must be in [1,1024]");
  const size_t staged = ((size_t)p.k + (size_t)p.n_lists * p.k_in) * sizeof(Hit) + (size_t)p.n_lists * sizeof(int);
  if (staged <= 200 * 1024) {
    COLTT_CUDA(cudaFuncSetAttribute(merge_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
    merge_topk_kernel<true><<<p.nq, 256, staged, stream>>>(p);
  } else {
    merge_topk_kernel<false><<<p.nq, 256, (size_t)p.k * sizeof(Hit), stream>>>(p);
  }
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
