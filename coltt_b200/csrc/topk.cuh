// topk.cuh — warp-level bounded top-K lists and the rank-based K-way merge.
//
// Replaces edge.PriorityQueue (edge/priority_queue.go:27-75 over
// edge/priorityqueue/priority_queue.go) and the shard-queue merge in VertexSearch
// (edge/none_vectorstore.go:173-178).  Where the reference pays a mutex, two heap
// allocations and an O(log K) sift for EVERY row, a warp here compares each score with the
// running K-th threshold held in a register and touches its list only for the ~K·ln(n/K)
// rows that actually enter it.
#pragma once
#include "common.cuh"

namespace coltt {

__device__ __forceinline__ Hit ld_hit(const Hit* p) {
  uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
  Hit h;
  h.id = ((uint64_t)v.y << 32) | v.x;
  h.score = __uint_as_float(v.z);
  h.slot = v.w;
  return h;
}
__device__ __forceinline__ void st_hit(Hit* p, const Hit& h) {
  uint4 v;
  v.x = (uint32_t)h.id;
  v.y = (uint32_t)(h.id >> 32);
  v.z = __float_as_uint(h.score);
  v.w = h.slot;
  __stcg(reinterpret_cast<uint4*>(p), v);
}

// Threshold pre-filter: false only if `score` is strictly worse than the K-th score (NaN on
// either side never rejects here; the exact (score,id) decision is taken in warp_list_insert).
__device__ __forceinline__ bool maybe_enters(float score, float kth_score, int nearest) {
  return !(nearest ? (score > kth_score) : (score < kth_score));
}

// Whole-warp insert of one candidate into a best-first sorted list of capacity k that lives in
// global memory (L2-resident; touched rarely).  `cnt` and `kth` are warp-uniform registers.
__device__ __forceinline__ void warp_list_insert(Hit* L, uint32_t k, uint32_t& cnt, float& kth, float score, uint32_t slot,
                                                 uint64_t id, int nearest) {
  const uint32_t lane = threadIdx.x & 31;
  // position = number of current entries that rank better than the candidate
  uint32_t pos = 0;
  for (uint32_t base = 0; base < cnt; base += 32) {
    uint32_t i = base + lane;
    bool b = false;
    if (i < cnt) {
      Hit e = ld_hit(L + i);
      b = better(e.score, e.id, score, id, nearest);
    }
    pos += __popc(__ballot_sync(0xffffffffu, b));
  }
  if (pos >= k) return;
  const uint32_t end = cnt < k ? cnt : k - 1;  // entries [pos, end) move to [pos+1, end]
  for (uint32_t hi = end; hi > pos;) {
    uint32_t span = hi - pos < 32 ? hi - pos : 32;
    Hit tmp;
    bool act = lane < span;
    uint32_t i = hi - lane;  // destination index
    if (act) tmp = ld_hit(L + i - 1);
    __syncwarp();
    if (act) st_hit(L + i, tmp);
    __syncwarp();
    hi -= span;
  }
  if (lane == 0) {
    Hit c;
    c.id = id;
    c.score = score;
    c.slot = slot;
    st_hit(L + pos, c);
  }
  if (cnt < k) cnt++;
  __syncwarp();
  if (cnt == k) kth = ld_hit(L + (k - 1)).score;
}

// number of entries of a best-first list that rank better than (score,id).  rev: the list is
// stored worst-first (a public T-order list seen from COLTT_COMPAT), entry i lives at n-1-i.
__device__ __forceinline__ uint32_t count_better(const Hit* L, uint32_t n, float score, uint64_t id, int nearest, int rev) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    Hit e = L[rev ? n - 1 - mid : mid];
    if (better(e.score, e.id, score, id, nearest)) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Rank-based merge of n_lists best-first lists staged in shared memory (L[j*k_in + i], cnt[j]):
// every candidate computes its global rank = own position + sum over the other lists of
// count_better(); ranks < k are scattered straight to sel[rank].  Exactly min(k,total)
// candidates have rank < k and each list contributes a prefix, so the work is
// ~(k + n_lists) candidates x n_lists binary searches, with no sort and no atomics.
// Ids must be unique across lists (they are: rows are partitioned).  Block-wide; sel in smem.
__device__ __forceinline__ void rank_merge_block(const Hit* L, const int* cnt, int n_lists, uint32_t k_in, uint32_t k,
                                                 int nearest, Hit* sel, int rev = 0, size_t list_stride = 0,
                                                 size_t cnt_stride = 1) {
  if (list_stride == 0) list_stride = k_in;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int j = warp; j < n_lists; j += n_warps) {
    uint32_t c = (uint32_t)cnt[(size_t)j * cnt_stride];
    if (c > k_in) c = k_in;
    for (uint32_t p0 = 0; p0 < c; p0 += 32) {
      const uint32_t pidx = p0 + lane;
      const bool active = pidx < c;
      uint32_t rank = k;
      if (active) {
        Hit cand = L[(size_t)j * list_stride + (rev ? c - 1 - pidx : pidx)];
        rank = pidx;
        for (int jj = 0; jj < n_lists && rank < k; jj++) {
          if (jj == j) continue;
          uint32_t cj = (uint32_t)cnt[(size_t)jj * cnt_stride];
          if (cj > k_in) cj = k_in;
          rank += count_better(L + (size_t)jj * list_stride, cj, cand.score, cand.id, nearest, rev);
        }
        if (rank < k) sel[rank] = cand;
      }
      if (__any_sync(0xffffffffu, !active || rank >= k)) break;  // the rest of this list ranks even lower
    }
  }
}

}  // namespace coltt
