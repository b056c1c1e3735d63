// topk_merge.cu — K5: final merge of per-CTA (or per-GPU) top-K lists into the answer.
//
// Replaces the tail of VertexSearch: re-Adding the 16 shard-local queues into the global
// queue and PriorityQueue.ToSlice's sort (edge/none_vectorstore.go:173-179,
// edge/priority_queue.go:57-69); across GPUs it is the one exchange step of the sharded search
// (after the all-gather of per-shard lists).  One CTA per query:
//   1. pivot: the k-th best of {the first ceil(k/n_lists) entries of every list} bounds the answer;
//   2. prune: each sorted list keeps only its prefix that is at least as good as the pivot
//      (binary search) — typically ~k..3k survivors out of n_lists*k candidates;
//   3. rank: all-pairs rank counting among the survivors scatters them straight into T order.
// No sort, no atomics on the data path; falls back to the generic rank merge (topk.cuh) if the
// survivors overflow shared memory.
#include "kernels.cuh"
#include "store.h"
#include "topk.cuh"

namespace coltt {

static constexpr int kMergeThreads = 256;

template <bool STAGED>
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(MergeParams p, uint32_t piv_per_list, uint32_t piv_cap, uint32_t surv_cap) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t P_s, M_s, overflow_s, has_tau_s;
  __shared__ Hit tau_s;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  if (p.n_active && p.q_base + q >= *p.n_active) return;
  const uint32_t q_out = p.q_map ? p.q_map[p.q_base + q] : q;
  const uint32_t ostride = p.out_stride ? p.out_stride : p.k;
  const int rev = (!p.in_best_first && !p.nearest) ? 1 : 0;  // public T-order lists are worst-first for COMPAT
  const size_t lstr = p.list_stride_hits ? p.list_stride_hits : (size_t)p.nq * p.k_in;
  const size_t cstr = p.count_stride ? p.count_stride : (size_t)p.nq;
  Hit* sel = reinterpret_cast<Hit*>(smem);            // [k]      (fallback path)
  Hit* piv = sel + p.k;                                // [piv_cap]
  Hit* surv = piv + piv_cap;                           // [surv_cap]
  const Hit* L;
  const int* cnt;
  size_t list_stride, cnt_stride;
  if (tid == 0) { P_s = 0; M_s = 0; overflow_s = 0; has_tau_s = 0; }
  if (STAGED) {
    Hit* Ls = surv + surv_cap;                         // [n_lists][k_in]
    int* cnt_s = reinterpret_cast<int*>(Ls + (size_t)p.n_lists * p.k_in);
    for (int j = tid; j < p.n_lists; j += blockDim.x) {
      int c = p.list_bases ? reinterpret_cast<const int*>(p.list_bases[j] + p.counts_off)[q] : p.counts[(size_t)j * cstr + q];
      cnt_s[j] = c > (int)p.k_in ? (int)p.k_in : c;
    }
    __syncthreads();
    for (uint32_t i = tid; i < (uint32_t)p.n_lists * p.k_in; i += blockDim.x) {
      uint32_t j = i / p.k_in, e = i - j * p.k_in;
      if ((int)e < cnt_s[j])
        Ls[i] = p.list_bases ? reinterpret_cast<const Hit*>(p.list_bases[j])[(size_t)q * p.k_in + e] : p.lists[(size_t)j * lstr + (size_t)q * p.k_in + e];
    }
    L = Ls; cnt = cnt_s; list_stride = p.k_in; cnt_stride = 1;
  } else {
    L = p.lists + (size_t)q * p.k_in; cnt = p.counts + q;
    list_stride = lstr; cnt_stride = cstr;
  }
  __syncthreads();
  auto count_of = [&](int j) -> uint32_t { uint32_t c = (uint32_t)cnt[(size_t)j * cnt_stride]; return c > p.k_in ? p.k_in : c; };
  auto entry = [&](int j, uint32_t i, uint32_t c) -> Hit { return L[(size_t)j * list_stride + (rev ? c - 1 - i : i)]; };

  // ---- 1. pivot subset and its k-th best
  uint32_t my_total = 0;
  for (int j = tid; j < p.n_lists; j += blockDim.x) {
    const uint32_t c = count_of(j);
    my_total += c;
    const uint32_t m = c < piv_per_list ? c : piv_per_list;
    if (m) {
      const uint32_t off = atomicAdd(&P_s, m);
      for (uint32_t i = 0; i < m; i++) piv[off + i] = entry(j, i, c);
    }
  }
  // block-wide total of candidates
  __shared__ uint32_t total_s;
  if (tid == 0) total_s = 0;
  __syncthreads();
  if (my_total) atomicAdd(&total_s, my_total);
  __syncthreads();
  const uint32_t P = P_s, total = total_s;
  const uint32_t n_out = total < p.k ? total : p.k;
  if (P >= p.k) {
    for (uint32_t e = tid; e < P; e += blockDim.x) {
      const Hit me = piv[e];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < P; j++) {
        const Hit o = piv[j];
        rank += better(o.score, o.id, me.score, me.id, p.nearest) ? 1u : 0u;
      }
      if (rank == p.k - 1) { tau_s = me; has_tau_s = 1; }
    }
  }
  __syncthreads();
  const bool has_tau = has_tau_s != 0;
  const Hit tau = tau_s;

  // ---- 2. prune every list to the prefix that is at least as good as the pivot
  for (int j = tid; j < p.n_lists; j += blockDim.x) {
    const uint32_t c = count_of(j);
    uint32_t len = c;
    if (has_tau) {  // first index whose entry is strictly worse than tau
      uint32_t lo = 0, hi = c;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const Hit e = entry(j, mid, c);
        if (!better(tau.score, tau.id, e.score, e.id, p.nearest)) lo = mid + 1;
        else hi = mid;
      }
      len = lo;
    }
    if (len) {
      const uint32_t off = atomicAdd(&M_s, len);
      if (off + len <= surv_cap) for (uint32_t i = 0; i < len; i++) surv[off + i] = entry(j, i, c);
      else overflow_s = 1;
    }
  }
  __syncthreads();
  Hit* out = p.out + (size_t)q_out * ostride;
  if (overflow_s) {
    rank_merge_block(L, cnt, p.n_lists, p.k_in, p.k, p.nearest, sel, rev, list_stride, cnt_stride);
    __syncthreads();
    for (uint32_t i = tid; i < n_out; i += blockDim.x) out[p.nearest ? i : n_out - 1 - i] = sel[i];
  } else {
    // ---- 3. rank among survivors; sel order is best-first, T order is the mirror for COMPAT
    const uint32_t M = M_s;
    for (uint32_t e = tid; e < M; e += blockDim.x) {
      const Hit me = surv[e];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < M; j++) {
        const Hit o = surv[j];
        rank += better(o.score, o.id, me.score, me.id, p.nearest) ? 1u : 0u;
      }
      if (rank < n_out) out[p.nearest ? rank : n_out - 1 - rank] = me;
    }
  }
  if (tid == 0) p.out_counts[q_out] = (int)n_out;
}

int launch_merge_topk(const MergeParams& p, cudaStream_t stream) {
  if (p.nq == 0) return COLTT_OK;
  if (p.k == 0 || p.k > 1024 || p.k_in == 0 || p.n_lists <= 0) return fail(COLTT_ERR_UNSUPPORTED, "merge: k must be in [1,1024]");
  const uint32_t piv_per_list = (p.k + p.n_lists - 1) / p.n_lists;
  const uint32_t piv_cap = piv_per_list * (uint32_t)p.n_lists;
  uint32_t surv_cap = 4 * p.k > 1024 ? 4 * p.k : 1024;
  const size_t base = ((size_t)p.k + piv_cap + surv_cap) * sizeof(Hit);
  const size_t staged = base + (size_t)p.n_lists * p.k_in * sizeof(Hit) + (size_t)p.n_lists * sizeof(int);
  if (p.list_bases && staged > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "merge: peer-memory lists need the staged variant");
  if (staged <= 200 * 1024) {
    { int arc = kernel_attrs(merge_topk_kernel<true>, staged); if (arc) return arc; }
    merge_topk_kernel<true><<<p.nq, kMergeThreads, staged, stream>>>(p, piv_per_list, piv_cap, surv_cap);
  } else {
    if (base > 200 * 1024) return fail(COLTT_ERR_UNSUPPORTED, "merge: too many lists for shared memory");
    { int arc = kernel_attrs(merge_topk_kernel<false>, base); if (arc) return arc; }
    merge_topk_kernel<false><<<p.nq, kMergeThreads, base, stream>>>(p, piv_per_list, piv_cap, surv_cap);
  }
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

}  // namespace coltt
