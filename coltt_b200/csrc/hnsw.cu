// hnsw.cu — K4: device-resident core/vectorindex HNSW search (+ the Commit-blob loader).
//
// Replaces, for search:  Hnsw.Load    core/vectorindex/hnsw_commit.go:164-278
//                        Hnsw.Search  core/vectorindex/hnsw.go:243-278
//                        greedyClosestNeighbor hnsw.go:320-343, searchLevel hnsw.go:345-389,
//                        selectNeighbors hnsw.go:391-397, PriorityQueue core/vectorindex/priority_queue.go
// Data layout in HBM: rows [n][row_stride] fp32 as stored by Insert (already normalized for
// cosine, hnsw.go:105-107), ||row||^2 [n] in AVX lane order, ids [n], level [n], and one CSR
// over (vertex, level): vbase[v] indexes edge_off, neighbours of v at level l are
// edge_nbr[edge_off[vbase[v]+l] .. edge_off[vbase[v]+l+1]) as SLOTS, sorted by neighbour id
// ascending — the reference iterates a Go map (random order); ascending id is the
// deterministic order this build defines, and the walk below is order-exact with respect to it.
//
// One CTA per query.  The traversal is latency/gather bound: per expansion the CTA gathers the
// (<= mMax0) unvisited neighbour rows straight from HBM (2 lanes per row, 4 AVX-lane chains each —
// the same exact arithmetic as flat_scan.cu, so every distance is bit-identical to the Go path),
// then thread 0 replays the reference's sequential heap logic (Go container/heap up/down, restated)
// over the batch.  lowerBound is frozen per expansion in the reference (hnsw.go:357), which is
// what makes the batch legal.  Throughput comes from many resident CTAs (queries) per SM.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "exact_math.cuh"
#include "hnsw.h"
#include "store.h"

namespace coltt {

static constexpr int kHnswThreads = 128;
static constexpr uint32_t kHnswChunkMax = 32;    // neighbour rows gathered and scored per pass (8 per warp)
static constexpr uint32_t kCandCapMin = 1024;    // candidate min-heap capacity (shared memory): first attempt
static constexpr uint32_t kCandCapMax = 8192;    //   ... and the re-run after an overflow
static constexpr uint32_t kNoSlot = 0xffffffffu;
static constexpr uint32_t kRingMaxStages = 4;

struct HnswParams {
  const uint8_t* rows; uint32_t row_stride; uint32_t dim; uint32_t q_stride;
  const float* row_norm2; const uint64_t* ids; const int32_t* level;
  const uint32_t* vbase; const uint32_t* edge_off; const uint32_t* edge_nbr;
  const uint32_t* nbr0; uint32_t nbr0_stride;   // level-0 adjacency at a fixed stride, kNoSlot-terminated
  uint32_t n; uint32_t entry; int metric;
  const float* queries; const float* q_norm2;   // prepared (normalized) queries [nq][q_stride]
  uint32_t nq; uint32_t k; uint32_t ef;
  uint32_t chunk_rows;                          // rows per gather pass (<= kHnswChunkMax)
  uint32_t rs;                                  // shared-memory row stride in bytes (== 32 mod 128: conflict-free LDS.64)
  uint32_t cand_cap;
  uint32_t* visited;                            // [nq][words] bitmap, zeroed by the caller
  uint32_t visited_words;
  Hit* out; int* out_counts; uint32_t out_stride;
  const uint32_t* q_map;                        // CTA -> query (re-runs of a subset), or null
  uint2* qlog; uint32_t log_cap;                // [nq][log_cap] queue operations of the fast path (replayed after a tie)
  unsigned long long* stats;                    // [0] distance evaluations, [1] expansions, [2] queue overflows, [3] queries finished on the literal heaps, [4..7] H1 / H2 / H3 / NaN events
  // RING variant only (experimental, COLTT_HNSW_RING): rows are staged through per-warp rings of dim-chunks
  uint32_t ring_cb;                             // bytes of a row per ring stage (multiple of 32)
  uint32_t ring_cs;                             // shared-memory stride of one staged chunk (== 32 mod 128)
  uint32_t ring_stages;                         // stages per warp (<= kRingMaxStages)
};

// Go container/heap (src/container/heap/heap.go: up / down), keyed on priority only —
// core/vectorindex/priority_queue.go:160-199: min queue Less = a<b, max queue Less = a>b.
// Entries are {priority bits, slot} pairs (one LDS.64 each).  up/down move a hole instead of swapping:
// the comparisons made and the final array are exactly those of Go's swap-based loops.
struct SmemHeap {
  uint2* e; uint32_t n; bool is_max;
  __device__ __forceinline__ bool less(float a, float b) const { return is_max ? a > b : a < b; }
  __device__ __forceinline__ float top() const { return __uint_as_float(e[0].x); }
  __device__ __forceinline__ void up(uint32_t j, uint2 x) {
    const float xp = __uint_as_float(x.x);
    while (j > 0) {
      const uint32_t i = (j - 1) / 2;
      const uint2 par = e[i];
      if (!less(xp, __uint_as_float(par.x))) break;
      e[j] = par;
      j = i;
    }
    e[j] = x;
  }
  __device__ __forceinline__ void down(uint32_t i, uint32_t m, uint2 x) {
    const float xp = __uint_as_float(x.x);
    for (;;) {
      const uint32_t j1 = 2 * i + 1;
      if (j1 >= m) break;
      uint2 c = e[j1];
      uint32_t j = j1;
      if (j1 + 1 < m) {
        const uint2 c2 = e[j1 + 1];
        if (less(__uint_as_float(c2.x), __uint_as_float(c.x))) { c = c2; j = j1 + 1; }
      }
      if (!less(__uint_as_float(c.x), xp)) break;
      e[i] = c;
      i = j;
    }
    e[i] = x;
  }
  __device__ __forceinline__ void push(float p, uint32_t s) { up(n, make_uint2(__float_as_uint(p), s)); n++; }
  // heap.Pop: Swap(0, n-1); down(0, n-1); remove the last
  __device__ __forceinline__ void pop(float& p, uint32_t& s) {
    const uint32_t m = n - 1;
    const uint2 t = e[0];
    if (m > 0) down(0, m, e[m]);
    p = __uint_as_float(t.x); s = t.y;
    n = m;
  }
};

// The result set of searchLevel as a sorted array spread over warp 0's registers (lane L owns entries
// [L*R, L*R+R), ascending priority, +inf padding).  While Go's heap layout cannot matter, the reference's two
// heaps reduce to ordered-set semantics: resultVertices = the ef smallest distances seen, and the candidates
// that can still be expanded are exactly the unexpanded members of that set — a candidate evicted from it has
// priority > lowerBound for the rest of the walk (lowerBound never grows once the set is full), so popping it can
// only end the loop (hnsw.go:359-361), which "no unexpanded member left" does as well.  The layout matters only
// where equal priorities meet an operation that has to choose between them; every queue operation is logged, so
// the literal heaps can be rebuilt at that point (see the kernel):
//   H1  candidateVertices.Pop() with two unexpanded members at the minimum  -> the literal candidate heap decides
//   H2  resultVertices.Pop() with the two largest members equal, or a NaN   -> the walk continues on the literal heaps
//   H3  equal priorities among the first k+1 members at the end             -> the literal result heap orders them
template <int R>
struct WarpSorted {
  float p[R > 0 ? R : 1];
  uint32_t s[R > 0 ? R : 1];                      // slot; bit 31 = already expanded
  uint32_t n;
  bool has_tie;                                   // an equal pair was inserted at some point (sticky)
  static constexpr uint32_t kExpanded = 0x80000000u;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int r = 0; r < R; r++) { p[r] = __int_as_float(0x7f800000); s[r] = kNoSlot; }
    n = 0;
    has_tie = false;
  }
  // entry n-1-back (back = 0: the largest member, resultVertices.Peek(); back = 1: the second largest)
  __device__ __forceinline__ float from_top(uint32_t back, uint32_t lane) const {
    const uint32_t lim = n - back;                // entries [0, lim) are candidates; the last of them is wanted
    float v = p[0];
#pragma unroll
    for (int r = 1; r < R; r++) v = lane * R + r < lim ? p[r] : v;
    return __shfl_sync(0xffffffffu, v, ((lim - 1) / R) & 31);
  }
  __device__ __forceinline__ float top(uint32_t lane) const { return from_top(0, lane); }
  // false = H2 (or NaN): nothing was changed, the caller must continue on the literal heaps
  __device__ __forceinline__ bool insert(float d, uint32_t slot, uint32_t ef, uint32_t lane, bool check = true) {
    uint32_t cnt = 0;
    bool eq = false;
#pragma unroll
    for (int r = 0; r < R; r++) { cnt += p[r] < d ? 1u : 0u; eq |= p[r] == d; }
    const float up_p = __shfl_up_sync(0xffffffffu, p[R - 1], 1);
    const uint32_t up_s = __shfl_up_sync(0xffffffffu, s[R - 1], 1);
    const uint32_t below = __ballot_sync(0xffffffffu, cnt == (uint32_t)R);   // lanes entirely below d: 0 .. full-1
    if (d != d) return false;
    if (__any_sync(0xffffffffu, eq)) has_tie = true;
    if (check && has_tie && n >= ef) {            // this insert evicts: are the two largest afterwards equal?
      const float m1 = top(lane);
      if (d == m1) return false;
      if (d < m1 && n >= 2 && from_top(1, lane) == m1) return false;   // d > m1: d itself is the unique largest and goes
    }
    const uint32_t full = __popc(below);
    if (lane > full) {                            // everything moves up by one
#pragma unroll
      for (int r = R - 1; r >= 1; r--) { p[r] = p[r - 1]; s[r] = s[r - 1]; }
      p[0] = up_p; s[0] = up_s;
    } else if (lane == full) {                    // d lands at this lane's entry `cnt`
#pragma unroll
      for (int r = R - 1; r >= 1; r--) {
        if ((uint32_t)r > cnt) { p[r] = p[r - 1]; s[r] = s[r - 1]; }
        else if ((uint32_t)r == cnt) { p[r] = d; s[r] = slot; }
      }
      if (cnt == 0) { p[0] = d; s[0] = slot; }
    }
    n++;
    if (n > ef) {                                  // resultVertices.Pop() of the largest (:379-381)
      n = ef;
#pragma unroll
      for (int r = 0; r < R; r++)
        if (lane * R + r >= ef) { p[r] = __int_as_float(0x7f800000); s[r] = kNoSlot; }
    }
    return true;
  }
  // drop the member with this slot (the literal result heap evicted it)
  __device__ __forceinline__ void remove(uint32_t slot, uint32_t lane) {
    uint32_t hit = R;
#pragma unroll
    for (int r = 0; r < R; r++)
      if ((s[r] & ~kExpanded) == slot && lane * R + r < n) hit = r;
    const uint32_t have = __ballot_sync(0xffffffffu, hit < (uint32_t)R);
    float dn_p = __shfl_down_sync(0xffffffffu, p[0], 1);
    uint32_t dn_s = __shfl_down_sync(0xffffffffu, s[0], 1);
    if (!have) return;
    if (lane == 31) { dn_p = __int_as_float(0x7f800000); dn_s = kNoSlot; }
    const uint32_t at = __ffs(have) - 1;
    if (lane >= at) {
#pragma unroll
      for (int r = 0; r < R - 1; r++)
        if (lane > at || (uint32_t)r >= hit) { p[r] = p[r + 1]; s[r] = s[r + 1]; }
      p[R - 1] = dn_p; s[R - 1] = dn_s;
    }
    n--;
  }
  // smallest unexpanded member -> (priority, slot), not marked; false if there is none
  __device__ __forceinline__ bool find(float& cp, uint32_t& cs, uint32_t lane) const {
    bool found = false;
    float vp = 0.0f;
    uint32_t vs = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const bool c = !found && !(s[r] & kExpanded) && lane * R + r < n;
      vp = c ? p[r] : vp;
      vs = c ? s[r] : vs;
      found |= c;
    }
    const uint32_t have = __ballot_sync(0xffffffffu, found);
    if (!have) return false;
    const uint32_t src = __ffs(have) - 1;
    cp = __shfl_sync(0xffffffffu, vp, src);
    cs = __shfl_sync(0xffffffffu, vs, src);
    return true;
  }
  // H1: more than one unexpanded member has priority cp
  __device__ __forceinline__ bool several(float cp, uint32_t lane) const {
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += (!(s[r] & kExpanded) && lane * R + r < n && p[r] == cp) ? 1u : 0u;
    const uint32_t some = __ballot_sync(0xffffffffu, c > 0);
    return __popc(some) > 1 || __any_sync(0xffffffffu, c > 1);
  }
  __device__ __forceinline__ void mark(uint32_t slot, uint32_t lane) {
#pragma unroll
    for (int r = 0; r < R; r++)
      if (s[r] == slot && lane * R + r < n) s[r] |= kExpanded;
  }
  // H3: equal neighbours among the first min(k+1, n) members
  __device__ __forceinline__ bool head_ties(uint32_t k, uint32_t lane) const {
    const uint32_t lim = k + 1 < n ? k + 1 : n;
    const float nxt0 = __shfl_down_sync(0xffffffffu, p[0], 1);
    bool t = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float nx = r + 1 < R ? p[r + 1 < R ? r + 1 : 0] : nxt0;
      t |= lane * R + r + 1 < lim && p[r] == nx;
    }
    return __any_sync(0xffffffffu, t);
  }
};

// One CTA per query.  Warp 0 drives the walk: it folds the pass that was just scored into the reference's
// sequential logic (the greedy scan; the result/candidate queues — WarpSorted while R > 0 and no tie was met,
// else lane 0 replaying Go's heaps) and decides the next pass; then it fetches that neighbour list,
// test-and-sets the visited bits, compacts the unvisited slots (ballot: list order is kept) and issues one
// cp.async.bulk (UBLKCP) per row, so every row of an expansion is in flight at once.  All four warps score the
// rows out of shared memory (4 lanes per row, 2 AVX-lane chains each: the exact avx.cpp arithmetic of
// flat_scan.cu).  lowerBound is frozen per expansion in the reference (hnsw.go:357), which is what makes the
// batch legal.  Two CTA barriers per pass; passes whose neighbours are all visited never leave warp 0.  Warp 0's
// control state is computed redundantly by all of its lanes (it depends only on shared and global memory).
enum { CMD_STOP = 0, CMD_ONE = 1, CMD_LIST = 2, CMD_L0 = 3 };
enum { ST_ENTRY = 0, ST_GREEDY = 1, ST_ENTRY2 = 2, ST_SEARCH = 3 };
enum { STOP_DONE = 1, STOP_OVERFLOW = 2, STOP_TIE = 3 };

// RING (experimental, off by default): instead of landing whole rows (32 x 3 KB per CTA at dim 768), every warp streams
// its own 8 rows through a small ring of dim-chunks, K1-style — accumulators stay in registers across chunks, so the
// arithmetic order is unchanged — which cuts the staging memory ~4x and lets four queries share an SM.
template <int METRIC, int R, bool RING = false>
__global__ void __launch_bounds__(kHnswThreads) hnsw_search_kernel(HnswParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* rows_s = smem;                                                        // [chunk_rows][rs]
  float* q_s = reinterpret_cast<float*>(smem + (size_t)p.chunk_rows * p.rs);    // [q_stride]
  uint2* cand_e = reinterpret_cast<uint2*>(q_s + p.q_stride);                    // [cand_cap]
  uint2* res_e = cand_e + p.cand_cap;                                            // [ef+1]
  uint32_t* nb_slot = reinterpret_cast<uint32_t*>(res_e + (p.ef + 1));          // [kHnswChunkMax]
  float* nb_dist = reinterpret_cast<float*>(nb_slot + kHnswChunkMax);            // [kHnswChunkMax]
  __shared__ uint32_t sh_cnt, sh_state;
  __shared__ __align__(8) uint64_t sh_bar;

  const uint32_t q = p.q_map ? p.q_map[blockIdx.x] : blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&sh_bar);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); sh_state = 0; sh_cnt = 0; }
  uint32_t ring_bar0 = 0, ring_phase = 0;          // RING: this warp's first stage barrier; one phase bit per stage
  if constexpr (RING) {
    __shared__ __align__(8) uint64_t sh_ring_bar[kHnswThreads / 32][kRingMaxStages];
    ring_bar0 = smem_u32(&sh_ring_bar[warp][0]);
    if (lane == 0) {
      for (uint32_t st = 0; st < kRingMaxStages; st++) mbar_init(ring_bar0 + 8 * st, 1);
      fence_mbar_init();
    }
  }
  for (uint32_t d = tid; d < p.q_stride; d += blockDim.x) q_s[d] = p.queries[(size_t)q * p.q_stride + d];
  __syncthreads();
  const float qn = METRIC == COLTT_COSINE ? p.q_norm2[q] : 0.0f;
  const uint32_t full8 = (p.dim / 8) * 8;
  const uint32_t cw = p.chunk_rows;
  uint32_t* vis = p.visited + (size_t)q * p.visited_words;
  uint32_t phase = 0;

  // warp 0's walk state
  unsigned long long evals = 0, exps = 0;
  SmemHeap cand{cand_e, 0, false}, res{res_e, 0, true};   // lane 0 only, once `literal`
  WarpSorted<R> ws;
  if constexpr (R > 0) ws.init();
  bool literal = R == 0;
  uint2* qlog = p.qlog + (size_t)q * p.log_cap;          // queue operations of the fast path, in order
  uint32_t log_n = 0;
  int stage = ST_ENTRY, l = 0;
  uint32_t ep = p.entry, best = kNoSlot, c0 = 0, e1 = 0, prev_m = 0, cur = 0, chunk = 0;
  bool need_list = true, chunk_more = false, started = false;
  float min_d = 0.0f, lb = 0.0f;
  // speculation: neighbour list and visited words of the likeliest next candidate, fetched under the current pass
  uint32_t spec_cur = kNoSlot, spec_s = kNoSlot, spec_w = 0;

  // lane 0: log records [0, cand_pos) / [0, res_pos) are in the literal candidate / result heap (brought up to date
  // only when Go's layout has to decide something)
  uint32_t cand_pos = 0, res_pos = 0;
  // an H2 eviction leaves a candidate outside the result set whose priority equals the largest member's: while that
  // holds it can still be popped and expanded (hnsw.go:359 tests `>`), so the literal candidate heap picks
  bool haz_set = false;
  float haz_x = 0.0f;
  // lane 0: one accepted neighbour through the reference's two heaps (hnsw.go:375-381); false = cand is full
  auto heap_accept = [&](float d, uint32_t slot) -> bool {
    if (cand.n >= p.cand_cap) return false;
    cand.push(d, slot);
    res.push(d, slot);
    if (res.n > p.ef) { float tp; uint32_t ts; res.pop(tp, ts); }
    return true;
  };
  // lane 0: replay log records [from, log_n) into Go's heaps (next record prefetched); false = cand is full
  auto replay = [&](uint32_t from, bool do_cand, bool do_res) -> bool {
    if (from >= log_n) return true;
    uint2 e = qlog[from];
    for (uint32_t i = from; i < log_n; i++) {
      const uint2 nx = i + 1 < log_n ? qlog[i + 1] : e;
      if (e.y == kNoSlot) {
        if (do_cand) { float tp; uint32_t ts; cand.pop(tp, ts); }
      } else {
        const float d = __uint_as_float(e.x);
        if (do_cand) {
          if (cand.n >= p.cand_cap) return false;
          cand.push(d, e.y);
        }
        if (do_res) {
          res.push(d, e.y);
          if (res.n > p.ef) { float tp; uint32_t ts; res.pop(tp, ts); }
        }
      }
      e = nx;
    }
    return true;
  };
  // leave the fast path for good: both literal heaps brought up to date; false = cand is full
  auto go_literal = [&]() -> bool {
    uint32_t ok = 1;
    if (lane == 0) {
      replay(res_pos, false, true);
      ok = replay(cand_pos, true, false) ? 1u : 0u;
      cand_pos = res_pos = log_n;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    literal = true;
    return ok != 0;
  };

  for (;;) {
    if (warp == 0) {
      uint32_t m = 0;
      for (;;) {
        uint32_t kind = CMD_STOP, a = 0, b = 0, stop = STOP_DONE;
        // ---- fold the pass that was just scored (nb_slot / nb_dist [0, prev_m)) into the walk
        if (!started) {
          started = true;                                     // nothing scored yet: first pass = the entrypoint (hnsw.go:253)
        } else if (stage == ST_ENTRY) {
          min_d = nb_dist[0];
          evals += 1;
          l = p.level[ep];
          stage = ST_GREEDY;
        } else if (stage == ST_GREEDY) {                      // greedyClosestNeighbor, hnsw.go:320-343
          for (uint32_t i = 0; i < prev_m; i++) {
            const float d = nb_dist[i];
            if (d < min_d) { min_d = d; best = nb_slot[i]; }
          }
          evals += prev_m;
        } else {                                              // ST_ENTRY2: hnsw.go:346-352; ST_SEARCH: hnsw.go:364-384
          evals += prev_m;
          const float dv = lane < prev_m ? nb_dist[lane] : 0.0f;
          const uint32_t sv = lane < prev_m ? nb_slot[lane] : 0u;
          uint32_t i0 = 0;                                    // first entry the literal heaps still have to see
          if constexpr (R > 0) {
            if (!literal) {
              const uint32_t all = prev_m >= 32 ? 0xffffffffu : (1u << prev_m) - 1u;
              const uint32_t lt = __ballot_sync(0xffffffffu, lane < prev_m && dv < lb) ;
              // (:374) `distance < lowerBound || len < ef`, entries in list order
              uint32_t todo = (stage == ST_ENTRY2 || ws.n < p.ef) ? all : lt;
              i0 = prev_m;
              while (todo) {
                const uint32_t i = __ffs(todo) - 1;
                todo &= todo - 1;
                const float d = __shfl_sync(0xffffffffu, dv, i);
                const uint32_t sl = __shfl_sync(0xffffffffu, sv, i);
                if (log_n >= p.log_cap) { stop = STOP_OVERFLOW; break; }
                if (!ws.insert(d, sl, p.ef, lane)) {
                  if (lane == 0) atomicAdd(p.stats + (d != d ? 7 : 5), 1ull);
                  if (d != d) {                               // NaN: this and the remaining entries go the literal way
                    i0 = i;
                    if (!go_literal()) stop = STOP_OVERFLOW;
                    break;
                  }
                  // H2: the literal result heap, brought up to date, takes the push and decides the eviction
                  uint32_t ts = 0;
                  if (lane == 0) {
                    replay(res_pos, false, true);
                    res.push(d, sl);
                    float tp;
                    res.pop(tp, ts);
                    res_pos = log_n + 1;                      // ... including the push logged below
                  }
                  ts = __shfl_sync(0xffffffffu, ts, 0);
                  haz_x = ws.top(lane);                       // priority of the evicted member == the largest one left
                  haz_set = true;
                  if (ts != sl) {
                    ws.remove(ts, lane);
                    ws.insert(d, sl, p.ef, lane, false);
                  }
                }
                if (lane == 0) {
                  qlog[log_n] = make_uint2(__float_as_uint(d), sl);
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.nbr0 + (size_t)sl * p.nbr0_stride));
                }
                log_n++;
                if (ws.n >= p.ef) todo &= lt;
              }
            }
          }
          if (literal && stop == STOP_DONE) {
            if (lane == 0) {
              for (uint32_t i = i0; i < prev_m; i++) {
                const float d = nb_dist[i];
                if (stage == ST_ENTRY2 || d < lb || res.n < p.ef) {   // (:374)
                  if (!heap_accept(d, nb_slot[i])) { stop = STOP_OVERFLOW; break; }
                }
              }
            }
            stop = __shfl_sync(0xffffffffu, stop, 0);
          }
          if (stage == ST_ENTRY2) {
            if (lane == 0) { atomicOr(vis + (ep >> 5), 1u << (ep & 31)); __threadfence_block(); }
            stage = ST_SEARCH;
          }
        }
        prev_m = 0;
        // ---- decide the next pass
        if (stop != STOP_DONE) {
          kind = CMD_STOP;
        } else if (stage == ST_ENTRY) {
          kind = CMD_ONE; a = ep;
        } else {
          if (stage == ST_GREEDY) {
            for (;;) {
              if (need_list) {
                if (l <= 0) { stage = ST_ENTRY2; break; }
                const uint32_t vb = p.vbase[ep];
                c0 = p.edge_off[vb + l]; e1 = p.edge_off[vb + l + 1];
                need_list = false;
              }
              if (c0 < e1) {
                kind = CMD_LIST; a = c0; b = e1 - c0 < cw ? e1 : c0 + cw;
                c0 = b;
                break;
              }
              if (best == kNoSlot) l--;                       // no strictly closer neighbour: next level down (:338-340)
              else { ep = best; best = kNoSlot; }             // move and scan again
              need_list = true;
            }
          }
          if (stage == ST_ENTRY2) {
            kind = CMD_ONE; a = ep;                           // entrypointDistance is recomputed (:346)
          } else if (stage == ST_SEARCH) {
            if (chunk_more) kind = CMD_L0;
            else {
              bool picked = false;
              if constexpr (R > 0) {
                if (!literal) {
                  picked = true;
                  float cp; uint32_t cs;
                  if (log_n >= p.log_cap) { stop = STOP_OVERFLOW; }
                  else {
                    // candidateVertices.Pop() (:355): the smallest unexpanded member — unless Go's layout has a say
                    bool found = ws.find(cp, cs, lane);
                    if (haz_set && ws.n && ws.top(lane) < haz_x) haz_set = false;
                    const bool consult = found ? (ws.has_tie && (ws.several(cp, lane) || (haz_set && cp == haz_x))) : haz_set;
                    if (consult) {                            // H1: the literal candidate heap, brought up to date, pops
                      uint32_t ok = 1, any = 0;
                      if (lane == 0) {
                        atomicAdd(p.stats + 4, 1ull);
                        ok = replay(cand_pos, true, false) ? 1u : 0u;
                        if (ok && cand.n) { cand.pop(cp, cs); any = 1; }
                        cand_pos = log_n + 1;                 // ... including the pop logged below
                      }
                      ok = __shfl_sync(0xffffffffu, ok, 0);
                      any = __shfl_sync(0xffffffffu, any, 0);
                      cp = __shfl_sync(0xffffffffu, cp, 0);
                      cs = __shfl_sync(0xffffffffu, cs, 0);
                      if (!ok) stop = STOP_OVERFLOW;
                      found = any && !(cp > ws.top(lane));    // (:354), (:359-361)
                    }
                    if (found && stop == STOP_DONE) {
                      ws.mark(cs, lane);
                      lb = ws.top(lane);                      // resultVertices.Peek() (:357)
                      cur = cs; kind = CMD_L0;
                      if (lane == 0) qlog[log_n] = make_uint2(0u, kNoSlot);
                      log_n++;
                    }
                  }
                }
              }
              if (!picked) {
                if (lane == 0) {
                  if (cand.n != 0) {                          // (:354)
                    float cp; uint32_t cs;
                    cand.pop(cp, cs);
                    lb = res.top();                           // resultVertices.Peek() (:357)
                    if (!(cp > lb)) { cur = cs; kind = CMD_L0; }   // (:359-361)
                  }
                }
                kind = __shfl_sync(0xffffffffu, kind, 0);
                cur = __shfl_sync(0xffffffffu, cur, 0);
                lb = __shfl_sync(0xffffffffu, lb, 0);
              }
              if (kind == CMD_L0) { chunk = 0; exps += 1; }
            }
            a = cur; b = chunk;
          }
        }
        if (kind == CMD_STOP) {
          if (lane == 0) sh_state = stop;
          m = 0;
          break;
        }
        uint32_t s = kNoSlot;
        bool valid = false;
        if (kind == CMD_ONE) { s = a; valid = lane == 0; }
        else if (kind == CMD_LIST) {
          valid = a + lane < b;
          if (valid) s = p.edge_nbr[a + lane];
        } else {
          // visited test-and-set for the whole chunk (:368-371), in parallel; the ballot below keeps list order
          const uint32_t off = b * cw + lane;
          const bool hit = a == spec_cur && b == 0;           // list and visited words are already here
          if (hit) s = spec_s;
          else if (lane < cw && off < p.nbr0_stride) s = p.nbr0[(size_t)a * p.nbr0_stride + off];
          const bool listed = s != kNoSlot;
          const uint32_t lm = __ballot_sync(0xffffffffu, listed);
          if (listed) {
            if (hit) {                                        // nothing touched this query's bitmap since spec_w was read
              valid = !((spec_w >> (s & 31)) & 1u);
              if (valid) atomicOr(vis + (s >> 5), 1u << (s & 31));
            } else {
              const uint32_t old = atomicOr(vis + (s >> 5), 1u << (s & 31));
              valid = !((old >> (s & 31)) & 1u);
            }
          }
          chunk_more = (uint32_t)__popc(lm) == cw && (b + 1) * cw < p.nbr0_stride;
          chunk = b + 1;
          spec_cur = kNoSlot;                                 // the bitmap just changed: spec_w is stale from here on
        }
        // compact the valid slots into nb_slot[] and start their row copies
        const uint32_t mask = __ballot_sync(0xffffffffu, valid);
        const uint32_t pos = __popc(mask & ((1u << lane) - 1u));
        m = __popc(mask);
        if (valid) nb_slot[pos] = s;
        if constexpr (!RING) {
          if (lane == 0 && m) mbar_arrive_expect_tx(bar, m * p.row_stride);
          __syncwarp();
          if (valid) bulk_g2s(smem_u32(rows_s + (size_t)pos * p.rs), p.rows + (size_t)s * p.row_stride, p.row_stride, bar);
        }
        if (m) break;
      }
      if (lane == 0) sh_cnt = m;
      prev_m = m;
      spec_cur = kNoSlot;
      if constexpr (R > 0) {
        if (m && !literal && stage == ST_SEARCH && !chunk_more && p.nbr0_stride <= cw) {
          float xp;
          if (ws.find(xp, spec_cur, lane)) spec_s = lane < p.nbr0_stride ? p.nbr0[(size_t)spec_cur * p.nbr0_stride + lane] : kNoSlot;
          else spec_cur = kNoSlot;
        }
      }
    }
    __syncthreads();
    if (sh_state) break;
    const uint32_t m = sh_cnt;
    // ---- all warps: exact distances of the m gathered rows -> nb_dist[]
    const uint32_t j = warp * 8 + (lane >> 2), h = lane & 3;
    const bool valid = j < m;
    float rn = 0.0f;
    if (METRIC == COLTT_COSINE && valid && h == 0) rn = p.row_norm2[nb_slot[j]];   // in flight while the rows land
    if constexpr (RING) {
      if (warp * 8 < m) {
        // this warp's rows [warp*8, warp*8 + nrows) stream through its own ring: stage = 8 rows x ring_cb bytes
        const uint32_t nrows = m - warp * 8 < 8u ? m - warp * 8 : 8u;
        const uint32_t CB = p.ring_cb, CS = p.ring_cs, S = p.ring_stages, CBE = CB / 4;
        const uint32_t n_chunks = (p.row_stride + CB - 1) / CB;
        uint8_t* ring = rows_s + (size_t)warp * S * 8 * CS;
        const uint8_t* src = p.rows + (size_t)(valid ? nb_slot[j] : 0u) * p.row_stride;
        auto issue_chunk = [&](uint32_t c) {
          const uint32_t st = c % S, off = c * CB;
          const uint32_t bytes = p.row_stride - off < CB ? p.row_stride - off : CB;
          if (lane == 0) mbar_arrive_expect_tx(ring_bar0 + 8 * st, bytes * nrows);
          __syncwarp();
          if (valid && h == 0) bulk_g2s(smem_u32(ring + ((size_t)st * 8 + (lane >> 2)) * CS), src + off, bytes, ring_bar0 + 8 * st);
        };
        for (uint32_t c = 0; c < S && c < n_chunks; c++) issue_chunk(c);
        float a0 = 0.0f, a1 = 0.0f, tot = 0.0f;     // AVX lanes 2h and 2h+1
        for (uint32_t c = 0; c < n_chunks; c++) {
          const uint32_t st = c % S;
          mbar_wait(ring_bar0 + 8 * st, (ring_phase >> st) & 1u);
          ring_phase ^= 1u << st;
          const uint8_t* rowp = ring + ((size_t)st * 8 + (valid ? (lane >> 2) : 0u)) * CS;
          const uint32_t e0 = c * CBE, e1 = e0 + CBE < full8 ? e0 + CBE : full8;
#pragma unroll 8
          for (uint32_t e = e0; e < e1; e += 8) {
            const float2 rv = *reinterpret_cast<const float2*>(rowp + (size_t)(e - e0 + 2 * h) * 4);
            const float2 qv = *reinterpret_cast<const float2*>(q_s + e + 2 * h);
            if (METRIC == COLTT_COSINE) {
              a0 = add_rn(a0, mul_rn(qv.x, rv.x)); a1 = add_rn(a1, mul_rn(qv.y, rv.y));
            } else {
              const float d0 = sub_rn(qv.x, rv.x), d1 = sub_rn(qv.y, rv.y);
              a0 = add_rn(a0, mul_rn(d0, d0)); a1 = add_rn(a1, mul_rn(d1, d1));
            }
          }
          if (c == n_chunks - 1) {                  // the scalar tail (dim % 8 elements) lives in the last chunk
            float t = add_rn(a0, a1);
            t = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
            tot = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
            for (uint32_t d = full8; d < p.dim; d++) {
              const float rv = reinterpret_cast<const float*>(rowp)[d - e0], qv = q_s[d];
              if (METRIC == COLTT_COSINE) tot = add_rn(tot, mul_rn(qv, rv));
              else { const float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
            }
          }
          __syncwarp();
          fence_proxy_async();                       // our reads of stage st precede its refill by the async proxy
          if (c + S < n_chunks) issue_chunk(c + S);
        }
        if (valid && h == 0) nb_dist[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, rn) : sqrt_via_f64(tot);
      }
    } else {
    mbar_wait(bar, phase);
    phase ^= 1;
    if (warp * 8 < m) {
      const uint8_t* rowp = rows_s + (size_t)(valid ? j : 0) * p.rs;
      float a0 = 0.0f, a1 = 0.0f;                 // AVX lanes 2h and 2h+1
#pragma unroll 8
      for (uint32_t e = 0; e < full8; e += 8) {
        const float2 rv = *reinterpret_cast<const float2*>(rowp + (size_t)(e + 2 * h) * 4);
        const float2 qv = *reinterpret_cast<const float2*>(q_s + e + 2 * h);
        if (METRIC == COLTT_COSINE) {
          a0 = add_rn(a0, mul_rn(qv.x, rv.x)); a1 = add_rn(a1, mul_rn(qv.y, rv.y));
        } else {
          const float d0 = sub_rn(qv.x, rv.x), d1 = sub_rn(qv.y, rv.y);
          a0 = add_rn(a0, mul_rn(d0, d0)); a1 = add_rn(a1, mul_rn(d1, d1));
        }
      }
      // ((l0+l1)+(l2+l3)) + ((l4+l5)+(l6+l7)): the reduction tree of flat_scan.cu (avx.cpp:4-9)
      float t = add_rn(a0, a1);
      t = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
      float tot = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
      for (uint32_t d = full8; d < p.dim; d++) {
        const float rv = reinterpret_cast<const float*>(rowp)[d], qv = q_s[d];
        if (METRIC == COLTT_COSINE) tot = add_rn(tot, mul_rn(qv, rv));
        else { const float df = sub_rn(qv, rv); tot = add_rn(tot, mul_rn(df, df)); }
      }
      if (valid && h == 0) nb_dist[j] = METRIC == COLTT_COSINE ? cosine_epilogue(tot, qn, rn) : sqrt_via_f64(tot);
    }
    }
    if (warp == 0 && spec_cur != kNoSlot) spec_w = spec_s != kNoSlot ? __ldcg(vis + (spec_s >> 5)) : 0u;
    __syncthreads();
  }
  // ---- selectNeighbors(k) (hnsw.go:391-397) and the back-to-front fill (:268-275)
  if (warp == 0) {
    const uint32_t state = sh_state;
    Hit* out = p.out + (size_t)q * p.out_stride;
    if (state != STOP_DONE) {
      if (lane == 0) {
        p.out_counts[q] = state == STOP_TIE ? -1 : 0;
        atomicAdd(p.stats + (state == STOP_TIE ? 3 : 2), 1ull);
      }
      return;                                      // the re-run counts this query's evaluations
    }
    bool done = false;
    if constexpr (R > 0) {
      if (!literal && ws.has_tie && ws.head_ties(p.k, lane)) {   // H3
        if (lane == 0) { atomicAdd(p.stats + 6, 1ull); replay(res_pos, false, true); }   // the result heap alone is needed
        literal = true;
      }
      if (!literal) {
        done = true;
        const uint32_t n_out = ws.n < p.k ? ws.n : p.k;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const uint32_t idx = lane * R + r;
          if (idx < n_out) {
            const uint32_t slot = ws.s[r] & ~WarpSorted<R>::kExpanded;
            Hit h; h.id = p.ids[slot]; h.score = ws.p[r]; h.slot = slot;
            out[idx] = h;
          }
        }
        if (lane == 0) p.out_counts[q] = (int)n_out;
      }
    }
    if (!done && lane == 0) {
      float tp; uint32_t ts;
      while (res.n > p.k) res.pop(tp, ts);
      const uint32_t n_out = res.n;
      for (int i = (int)n_out - 1; i >= 0; i--) {
        res.pop(tp, ts);
        Hit h; h.id = p.ids[ts]; h.score = tp; h.slot = ts;
        out[i] = h;
      }
      p.out_counts[q] = (int)n_out;
    }
    if (lane == 0) {
      atomicAdd(p.stats + 0, evals);
      atomicAdd(p.stats + 1, exps);
      if (R > 0 && literal) atomicAdd(p.stats + 3, 1ull);
    }
  }
}

// ------------------------------------------------------------------------------------------

struct BlobR {
  const uint8_t* p; size_t n, pos = 0; bool ok = true;
  uint64_t be(int nb) {
    if (pos + nb > n) { ok = false; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v = (v << 8) | p[pos++];
    return v;
  }
  void skip(size_t k) { if (pos + k > n) ok = false; else pos += k; }
};

// CSR over (vertex, level) from per-list edge vectors: neighbours sorted by id (the deterministic iteration
// order), duplicates dropped; edge distances ride along for Commit.
int hnsw_install_graph(Hnsw* h, const std::vector<uint32_t>& vbase, std::vector<std::vector<HnswEdge>>& lists) {
  const uint32_t n = h->n;
  std::vector<uint32_t> edge_off(vbase[n] + 1, 0), edge_nbr;
  std::vector<uint32_t> edge_dist;
  for (uint32_t i = 0; i < vbase[n]; i++) {
    auto& lst = lists[i];
    std::sort(lst.begin(), lst.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id < b.id; });
    lst.erase(std::unique(lst.begin(), lst.end(), [](const HnswEdge& a, const HnswEdge& b) { return a.id == b.id; }), lst.end());
    edge_off[i] = (uint32_t)edge_nbr.size();
    for (auto& e : lst) { edge_nbr.push_back(e.slot); edge_dist.push_back(e.dist_bits); }
  }
  edge_off[vbase[n]] = (uint32_t)edge_nbr.size();
  int rc;
  // level-0 lists again at a fixed stride (one dependent load per expansion instead of vbase -> edge_off -> edge_nbr)
  uint32_t deg0 = 0;
  for (uint32_t v = 0; v < n; v++) deg0 = std::max(deg0, edge_off[vbase[v] + 1] - edge_off[vbase[v]]);
  const uint32_t stride0 = std::max(32u, (deg0 + 31) / 32 * 32);
  std::vector<uint32_t> nbr0((size_t)n * stride0, 0xffffffffu);
  for (uint32_t v = 0; v < n; v++)
    std::copy(edge_nbr.begin() + edge_off[vbase[v]], edge_nbr.begin() + edge_off[vbase[v] + 1], nbr0.begin() + (size_t)v * stride0);
  for (void* ptr : {(void*)h->d_vbase, (void*)h->d_edge_off, (void*)h->d_edge_nbr, (void*)h->d_edge_dist, (void*)h->d_nbr0})
    if (ptr) cudaFree(ptr);
  h->d_vbase = h->d_edge_off = h->d_edge_nbr = h->d_edge_dist = h->d_nbr0 = nullptr;
  if ((rc = upload(&h->d_vbase, vbase)) || (rc = upload(&h->d_edge_off, edge_off)) || (rc = upload(&h->d_edge_nbr, edge_nbr)) ||
      (rc = upload(&h->d_edge_dist, edge_dist)) || (rc = upload(&h->d_nbr0, nbr0)))
    return rc;
  h->nbr0_stride = stride0;
  h->n_edges = edge_nbr.size();
  return COLTT_OK;
}

// Hnsw.Load(header=true): hnsw_commit.go:164-278 with hnsw_config.go:203-245 (config) and
// metadata.go:43-105 (per-vertex metadata records are skipped by length).
int hnsw_load(const void* blob, size_t len, int device, Hnsw** out) {
  int rc = require_device(device);
  if (rc) return rc;
  COLTT_CUDA(cudaSetDevice(device));
  BlobR r{(const uint8_t*)blob, len};
  std::unique_ptr<Hnsw> h(new Hnsw());
  h->device = device;
  h->search_algo = (int32_t)r.be(4);
  h->level_mult_bits = (uint32_t)r.be(4);  // levelMultiplier (insert-time only; kept for Commit)
  h->ef_default = (int32_t)r.be(4);
  h->ef_construction = (int32_t)r.be(4);
  h->m = (int32_t)r.be(4); h->m_max = (int32_t)r.be(4); h->m_max0 = (int32_t)r.be(4);
  h->dim = (uint32_t)r.be(4);
  const uint8_t di = (uint8_t)r.be(1);
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated HNSW commit header");
  if (di != 1 && di != 2) return fail(COLTT_ERR_FORMAT, "Invalid space type");   // InvalidSpaceTypeErr, hnsw_commit.go:33
  if (h->dim == 0) return fail(COLTT_ERR_FORMAT, "zero dimension in HNSW commit header");
  h->metric = di == 1 ? COLTT_COSINE : COLTT_EUCLIDEAN;
  h->row_stride = (h->dim * 4 + 15) / 16 * 16;
  cudaDeviceProp pr;
  COLTT_CUDA(cudaGetDeviceProperties(&pr, device));
  h->n_sms = pr.multiProcessorCount;
  COLTT_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  COLTT_CUDA(cudaMalloc((void**)&h->d_stats, 8 * sizeof(unsigned long long)));
  std::vector<uint64_t> ids;
  std::vector<int32_t> levels;
  std::vector<uint8_t> rows;
  std::unordered_map<uint64_t, uint32_t> id2slot;
  std::vector<std::vector<uint32_t>> shard_slots(16);
  uint64_t ep_id = 0;
  if (r.pos < len) {  // non-empty index
    ep_id = r.be(8);
    for (int sh = 0; sh < 16 && r.ok; sh++) {
      const uint32_t cnt = (uint32_t)r.be(4);
      for (uint32_t i = 0; i < cnt && r.ok; i++) {
        const uint64_t id = r.be(8);
        const int32_t lvl = (int32_t)r.be(4);
        if (!r.ok || lvl < 0 || lvl > 64) { r.ok = false; break; }
        if (r.pos + (size_t)h->dim * 4 > r.n) { r.ok = false; break; }
        const size_t off = rows.size();
        rows.resize(off + h->row_stride, 0);
        const uint8_t* src = r.p + r.pos;
        for (uint32_t d = 0; d < h->dim; d++)
          for (int b = 0; b < 4; b++) rows[off + (size_t)d * 4 + b] = src[(size_t)d * 4 + (3 - b)];
        r.skip((size_t)h->dim * 4);
        const uint32_t mc = (uint32_t)r.be(2);
        for (uint32_t k = 0; k < mc && r.ok; k++) { r.skip(r.be(1)); r.skip(r.be(2)); }
        const uint32_t slot = (uint32_t)ids.size();
        id2slot[id] = slot;
        ids.push_back(id);
        levels.push_back(lvl);
        shard_slots[sh].push_back(slot);
      }
    }
  }
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated HNSW commit blob (vertices)");
  const uint32_t n = (uint32_t)ids.size();
  std::vector<uint32_t> vbase(n + 1, 0);
  for (uint32_t v = 0; v < n; v++) vbase[v + 1] = vbase[v] + (uint32_t)levels[v] + 1;
  std::vector<std::vector<HnswEdge>> lists(vbase[n]);
  for (int sh = 0; sh < 16 && r.ok; sh++)
    for (size_t i = 0; i < shard_slots[sh].size() && r.ok; i++) {
      const uint64_t id = r.be(8);
      auto it = id2slot.find(id);
      if (!r.ok || it == id2slot.end()) { r.ok = false; break; }
      const uint32_t v = it->second;
      for (int l = levels[v]; l >= 0 && r.ok; l--) {
        const uint32_t ne = (uint32_t)r.be(4);
        auto& lst = lists[vbase[v] + l];
        for (uint32_t j = 0; j < ne && r.ok; j++) {
          const uint64_t nid = r.be(8);
          const uint32_t dbits = (uint32_t)r.be(4);   // stored edge distance: not needed by Search, kept for Commit
          auto nt = id2slot.find(nid);
          if (nt == id2slot.end()) { r.ok = false; break; }
          lst.push_back(HnswEdge{nid, nt->second, dbits});
        }
      }
    }
  if (!r.ok) return fail(COLTT_ERR_FORMAT, "truncated or inconsistent HNSW commit blob (edges)");
  h->n = n;
  for (int32_t l : levels) h->max_level = std::max(h->max_level, l);
  if (n) {
    auto it = id2slot.find(ep_id);
    if (it == id2slot.end()) return fail(COLTT_ERR_FORMAT, "entrypoint id not among the vertices");
    h->entry = it->second;
  }
  if ((rc = upload(&h->d_rows, rows)) || (rc = upload(&h->d_ids, ids)) || (rc = upload(&h->d_level, levels))) return rc;
  if ((rc = hnsw_install_graph(h.get(), vbase, lists))) return rc;
  COLTT_CUDA(cudaMalloc((void**)&h->d_norm2, std::max<size_t>(n, 1) * 4));
  if (n) {
    rc = launch_norm2_stored_f32(h->d_rows, h->row_stride, h->dim, n, h->d_norm2, h->stream);
    if (rc) return rc;
    COLTT_CUDA(cudaStreamSynchronize(h->stream));
  }
  *out = h.release();
  return COLTT_OK;
}


static int launch_hnsw_search(int metric, int R, const HnswParams& p, unsigned n_ctas, size_t smem, cudaStream_t st) {
  if (p.ring_cb) {   // dim-chunk ring staging (the default for wide rows), register result set only
#define COLTT_HNSW_RING_CASE(M, RR)                                                          \
  if (metric == M && R == RR) {                                                              \
    int arc = kernel_attrs(hnsw_search_kernel<M, RR, true>, smem);                           \
    if (arc) return arc;                                                                     \
    hnsw_search_kernel<M, RR, true><<<n_ctas, kHnswThreads, smem, st>>>(p);                  \
  } else
    COLTT_HNSW_RING_CASE(COLTT_COSINE, 2) COLTT_HNSW_RING_CASE(COLTT_COSINE, 4) COLTT_HNSW_RING_CASE(COLTT_COSINE, 8)
    COLTT_HNSW_RING_CASE(COLTT_EUCLIDEAN, 2) COLTT_HNSW_RING_CASE(COLTT_EUCLIDEAN, 4) COLTT_HNSW_RING_CASE(COLTT_EUCLIDEAN, 8)
    return fail(COLTT_ERR_UNSUPPORTED, "ring staging needs the register result set (ef <= 256)");
#undef COLTT_HNSW_RING_CASE
    count_launch();
    COLTT_CUDA(cudaGetLastError());
    return COLTT_OK;
  }
#define COLTT_HNSW_CASE(M, RR)                                                               \
  if (metric == M && R == RR) {                                                              \
    int arc = kernel_attrs(hnsw_search_kernel<M, RR>, smem);                                 \
    if (arc) return arc;                                                                     \
    hnsw_search_kernel<M, RR><<<n_ctas, kHnswThreads, smem, st>>>(p);                        \
  } else
  COLTT_HNSW_CASE(COLTT_COSINE, 0) COLTT_HNSW_CASE(COLTT_COSINE, 2) COLTT_HNSW_CASE(COLTT_COSINE, 4) COLTT_HNSW_CASE(COLTT_COSINE, 8)
  COLTT_HNSW_CASE(COLTT_EUCLIDEAN, 0) COLTT_HNSW_CASE(COLTT_EUCLIDEAN, 2) COLTT_HNSW_CASE(COLTT_EUCLIDEAN, 4) COLTT_HNSW_CASE(COLTT_EUCLIDEAN, 8)
  return fail(COLTT_ERR_UNSUPPORTED, "no HNSW kernel for this metric / result-set width");
#undef COLTT_HNSW_CASE
  count_launch();
  COLTT_CUDA(cudaGetLastError());
  return COLTT_OK;
}

static int hnsw_search(Hnsw* h, const float* queries, size_t nq, int k, int ef_in, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || !out_ids || !out_scores || !out_counts) return fail(COLTT_ERR_INVALID, "null argument");
  if (k <= 0) return fail(COLTT_ERR_INVALID, "k must be positive");
  std::lock_guard<std::mutex> lk(h->mu);
  COLTT_CUDA(cudaSetDevice(h->device));
  if (h->n == 0) {  // `if entrypoint == nil { return make(SearchResult, 0), nil }` hnsw.go:249-251
    for (size_t q = 0; q < nq; q++) out_counts[q] = 0;
    return COLTT_OK;
  }
  const uint32_t ef = (uint32_t)std::max(ef_in > 0 ? ef_in : h->ef_default, k);   // gomath.MaxInt(ef, k) hnsw.go:258
  const uint32_t q_stride = (h->dim + 7) / 8 * 8;
  // Shared memory per CTA = gathered rows + query + both heaps.  Two CTAs per SM when the batch has more queries
  // than SMs (the walk is latency bound: a second resident query hides the first one's serial heap work).
  const uint32_t rs = (h->row_stride + 127) / 128 * 128 + 32;
  const size_t sm_total = 227 * 1024;
  // R > 0: result set in warp registers (ef <= 32*R); R == 0: literal Go-heap replay in shared memory (large ef, and
  // the re-run of queries whose walk met equal priorities).
  static const int env_chunk = getenv("COLTT_HNSW_CHUNK") ? atoi(getenv("COLTT_HNSW_CHUNK")) : 0;
  static const int env_ctas = getenv("COLTT_HNSW_CTAS") ? atoi(getenv("COLTT_HNSW_CTAS")) : 0;
  static const int env_literal = getenv("COLTT_HNSW_LITERAL") ? atoi(getenv("COLTT_HNSW_LITERAL")) : 0;
  // Rows are staged through per-warp rings of dim-chunks ("<chunk bytes>,<stages>", default 512,2 for rows of 1 KB and more;
  // COLTT_HNSW_RING=0 selects whole-row staging): a CTA then needs ~50 KB of shared memory instead of ~100 KB at dim 768, four
  // queries are resident per SM instead of two, and the walk — latency bound — gains 23 % (profiles/r2_hnsw_summary.md).
  static const char* env_ring_raw = getenv("COLTT_HNSW_RING");
  const char* env_ring = env_ring_raw ? env_ring_raw : (h->row_stride >= 1024 ? "512,2" : nullptr);
  int R = env_literal ? 0 : ef <= 64 ? 2 : ef <= 128 ? 4 : ef <= 256 ? 8 : 0;
  uint32_t cand_cap = kCandCapMin;
  while (cand_cap < 8 * ef && cand_cap < kCandCapMax) cand_cap *= 2;
  uint32_t chunk_rows = 0;
  size_t smem = 0;
  auto plan = [&](uint32_t cap, size_t n_ctas) -> bool {
    const size_t fixed = (size_t)q_stride * 4 + (size_t)cap * 8 + (size_t)(ef + 1) * 8 + kHnswChunkMax * 8;
    int ctas = env_ctas > 0 ? env_ctas : (n_ctas > (size_t)h->n_sms ? 2 : 1);
    for (; ctas >= 1; ctas--) {
      const size_t budget = sm_total / ctas - 1024 - 128;     // 1 KB reserved per CTA + the static shared variables
      if (budget <= fixed) continue;
      uint32_t c = (uint32_t)std::min<size_t>((budget - fixed) / rs, kHnswChunkMax);
      if (env_chunk > 0) c = std::min<uint32_t>(c, (uint32_t)env_chunk);
      if (c >= 8 || (ctas == 1 && c >= 1)) {
        chunk_rows = c;
        smem = fixed + (size_t)c * rs;
        return true;
      }
    }
    return false;
  };
  if (!plan(cand_cap, nq)) return fail(COLTT_ERR_UNSUPPORTED, "ef/dim too large for the HNSW kernel's shared memory");
  uint32_t ring_cb = 0, ring_cs = 0, ring_stages = 0;
  if (env_ring && R > 0) {
    unsigned cb = 0, stg = 0;
    if (sscanf(env_ring, "%u,%u", &cb, &stg) == 2 && cb >= 32 && cb % 32 == 0 && stg >= 2 && stg <= kRingMaxStages) {
      const uint32_t r_cb = std::min<uint32_t>(cb, (h->row_stride + 31) / 32 * 32);
      const uint32_t r_cs = (r_cb + 127) / 128 * 128 + 32;
      const size_t fixed = (size_t)q_stride * 4 + (size_t)cand_cap * 8 + (size_t)(ef + 1) * 8 + kHnswChunkMax * 8;
      const size_t r_smem = fixed + (size_t)kHnswChunkMax * stg * r_cs;
      if (r_smem <= 200 * 1024) {
        ring_cb = r_cb; ring_cs = r_cs; ring_stages = stg;
        chunk_rows = kHnswChunkMax;               // the rows region is chunk_rows * rs bytes: 32 rows x stages x ring_cs
        smem = r_smem;
      }
    }
  }
  const uint32_t words = (h->n + 31) / 32;
  cudaStream_t st = h->stream;
  int rc;
  if ((rc = h->q_in.ensure(nq * h->dim * 4)) || (rc = h->q_deq.ensure(nq * q_stride * 4)) || (rc = h->q_n2.ensure(nq * 4)) ||
      (rc = h->visited.ensure(nq * (size_t)words * 4)) || (rc = h->out.ensure(nq * (size_t)k * sizeof(Hit))) || (rc = h->counts.ensure(nq * 4)))
    return rc;
  COLTT_CUDA(cudaMemcpyAsync(h->q_in.p, queries, nq * h->dim * 4, cudaMemcpyHostToDevice, st));
  PrepParams pp{};
  pp.in = (const float*)h->q_in.p; pp.n = nq; pp.in_stride = h->dim; pp.dim = h->dim; pp.smem_stride = (h->dim + 3) / 4 * 4;
  pp.normalize = h->metric == COLTT_COSINE;   // hnsw.go:244-246
  pp.norm2_out = (float*)h->q_n2.p; pp.deq_out = (float*)h->q_deq.p; pp.deq_stride = q_stride;
  rc = launch_prep_rows(pp, ELEM_F32, st);
  if (rc) return rc;
  HnswParams p{};
  p.rows = h->d_rows; p.row_stride = h->row_stride; p.dim = h->dim; p.q_stride = q_stride; p.row_norm2 = h->d_norm2; p.ids = h->d_ids;
  p.level = h->d_level; p.vbase = h->d_vbase; p.edge_off = h->d_edge_off; p.edge_nbr = h->d_edge_nbr; p.n = h->n; p.entry = h->entry;
  p.nbr0 = h->d_nbr0; p.nbr0_stride = h->nbr0_stride;
  p.metric = h->metric; p.queries = (const float*)h->q_deq.p; p.q_norm2 = (const float*)h->q_n2.p; p.nq = (uint32_t)nq; p.k = (uint32_t)k;
  p.ef = ef; p.visited = (uint32_t*)h->visited.p; p.visited_words = words; p.out = (Hit*)h->out.p; p.out_counts = (int*)h->counts.p;
  p.out_stride = (uint32_t)k; p.stats = h->d_stats; p.rs = ring_cb ? ring_stages * ring_cs : rs;
  p.ring_cb = ring_cb; p.ring_cs = ring_cs; p.ring_stages = ring_stages;
  std::vector<Hit> hits(nq * (size_t)k);
  unsigned long long stats[8], evals = 0, exps = 0, ties = 0;
  float kernel_ms = 0.0f;
  std::vector<uint32_t> redo;                     // queries to run again (empty = the whole batch)
  for (;;) {
    const size_t n_ctas = redo.empty() ? nq : redo.size();
    p.chunk_rows = chunk_rows; p.cand_cap = cand_cap; p.log_cap = 2 * cand_cap;
    if ((rc = h->qlog.ensure(R > 0 ? nq * (size_t)p.log_cap * 8 : 8))) return rc;
    p.qlog = (uint2*)h->qlog.p;
    p.q_map = nullptr;
    if (!redo.empty()) {
      if ((rc = h->q_map.ensure(redo.size() * 4))) return rc;
      COLTT_CUDA(cudaMemcpyAsync(h->q_map.p, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
      p.q_map = (const uint32_t*)h->q_map.p;
    }
    COLTT_CUDA(cudaMemsetAsync(h->visited.p, 0, nq * (size_t)words * 4, st));
    COLTT_CUDA(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(unsigned long long), st));
    if (!h->ev0) { COLTT_CUDA(cudaEventCreate(&h->ev0)); COLTT_CUDA(cudaEventCreate(&h->ev1)); }
    COLTT_CUDA(cudaEventRecord(h->ev0, st));
    if ((rc = launch_hnsw_search(h->metric, R, p, (unsigned)n_ctas, smem, st))) return rc;
    COLTT_CUDA(cudaEventRecord(h->ev1, st));
    COLTT_CUDA(cudaMemcpyAsync(hits.data(), h->out.p, hits.size() * sizeof(Hit), cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaMemcpyAsync(out_counts, h->counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaMemcpyAsync(stats, h->d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
    COLTT_CUDA(cudaStreamSynchronize(st));
    evals += stats[0];
    exps += stats[1];
    { float ms = 0.0f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); kernel_ms += ms; }
    static const bool dbg = getenv("COLTT_HNSW_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[hnsw] pass: ctas=%zu R=%d chunk=%u smem=%zu cap=%u overflows=%llu literal=%llu H1=%llu H2=%llu H3=%llu NaN=%llu\n", n_ctas, R, chunk_rows, smem, cand_cap, stats[2], stats[3], stats[4], stats[5], stats[6], stats[7]);
    ties += stats[3];
    if (!stats[2]) break;
    // a candidate queue (or the fast path's operation log) outgrew its capacity: those queries run again on the
    // literal heaps with the largest one
    if ((R == 0 && cand_cap >= kCandCapMax) || !plan(kCandCapMax, stats[2]))
      return fail(COLTT_ERR_UNSUPPORTED, "HNSW candidate queue overflowed its shared-memory capacity (ef too large)");
    R = 0;
    cand_cap = kCandCapMax;
    p.ring_cb = 0; p.rs = rs;                      // the literal re-run uses the whole-row staging
    std::vector<uint32_t> again;
    if (redo.empty()) {
      for (size_t q = 0; q < nq; q++)
        if (out_counts[q] <= 0) again.push_back((uint32_t)q);
    } else {
      for (uint32_t q : redo)
        if (out_counts[q] <= 0) again.push_back(q);
    }
    if (again.empty()) return fail(COLTT_ERR_UNSUPPORTED, "HNSW re-run bookkeeping lost its queries");
    redo.swap(again);
  }
  // HnswSearchHeuristic (hnsw.go:262-266,399-447) with the defaults a loaded index gets (extendCandidates = false,
  // keepPruned = true; the Commit header does not carry them, hnsw_config.go:179-245) keeps the k smallest of the ef
  // results exactly like selectNeighbors — except among EQUAL priorities, where Go's two heap walks may keep different
  // vertices.  The walk above replays selectNeighbors' heaps; rather than answer a heuristic-configured index with
  // simple-mode tie-breaking, fail loudly when a query of this batch actually met equal priorities.
  if (h->search_algo == 1 && ties > 0)
    return fail(COLTT_ERR_UNSUPPORTED, "HnswSearchHeuristic: this batch met equal priorities, whose order under the heuristic selector is not reproduced");
  stats[0] = evals;
  stats[1] = exps;
  h->last_ties = ties;
  h->last_kernel_ms = kernel_ms;
  h->last_evals = stats[0];
  h->last_exp = stats[1];
  for (size_t q = 0; q < nq; q++)
    for (int i = 0; i < out_counts[q]; i++) {
      out_ids[q * (size_t)k + i] = hits[q * (size_t)k + i].id;
      out_scores[q * (size_t)k + i] = hits[q * (size_t)k + i].score;
    }
  return COLTT_OK;
}

// Hnsw.Search whose per-query hits also stay in device memory (the sharded search's send side, comm.cu).  The pointers
// are the handle's own scratch: valid until the next search on this handle.
int hnsw_search_keep_device(Hnsw* h, const float* queries, size_t nq, int k, int ef, const Hit** d_hits, const int** d_counts, cudaStream_t* st) {
  std::vector<uint64_t> ids(nq * (size_t)k);
  std::vector<float> sc(nq * (size_t)k);
  std::vector<int32_t> cnt(nq);
  int rc = hnsw_search(h, queries, nq, k, ef, ids.data(), sc.data(), cnt.data());
  if (rc) return rc;
  if (h->n == 0) {   // empty sub-graph: hnsw_search answered on the host, the device scratch may not exist yet
    std::lock_guard<std::mutex> lk(h->mu);
    COLTT_CUDA(cudaSetDevice(h->device));
    if ((rc = h->out.ensure(nq * (size_t)k * sizeof(Hit))) || (rc = h->counts.ensure(nq * 4))) return rc;
    COLTT_CUDA(cudaMemset(h->counts.p, 0, nq * 4));
  }
  *d_hits = (const Hit*)h->out.p;
  *d_counts = (const int*)h->counts.p;
  *st = h->stream;
  return COLTT_OK;
}

}  // namespace coltt

using coltt::fail;
using coltt::Hnsw;

extern "C" {
COLTT_API int coltt_b200_hnsw_load(const void* commit_blob, size_t len, int device, coltt_hnsw** out) {
  if (!commit_blob || !out) return fail(COLTT_ERR_INVALID, "null argument");
  Hnsw* h = nullptr;
  int rc = coltt::hnsw_load(commit_blob, len, device, &h);
  if (rc == COLTT_OK) *out = reinterpret_cast<coltt_hnsw*>(h);
  return rc;
}
COLTT_API void coltt_b200_hnsw_destroy(coltt_hnsw* h) { delete reinterpret_cast<Hnsw*>(h); }
COLTT_API int coltt_b200_hnsw_dim(coltt_hnsw* h, uint32_t* dim) {
  if (!h || !dim) return fail(COLTT_ERR_INVALID, "null argument");
  *dim = reinterpret_cast<Hnsw*>(h)->dim;
  return COLTT_OK;
}
COLTT_API int coltt_b200_hnsw_len(coltt_hnsw* h, uint64_t* n) {
  if (!h || !n) return fail(COLTT_ERR_INVALID, "null argument");
  *n = reinterpret_cast<Hnsw*>(h)->n;
  return COLTT_OK;
}
COLTT_API int coltt_b200_hnsw_search(coltt_hnsw* h, const float* queries, size_t nq, int k, int ef, uint64_t* out_ids, float* out_scores,
                                     int32_t* out_counts) {
  if (!h) return fail(COLTT_ERR_INVALID, "null index");
  return coltt::hnsw_search(reinterpret_cast<Hnsw*>(h), queries, nq, k, ef, out_ids, out_scores, out_counts);
}
COLTT_API int coltt_b200_hnsw_last_timing(coltt_hnsw* h, float* kernel_ms) {
  if (!h || !kernel_ms) return fail(COLTT_ERR_INVALID, "null argument");
  *kernel_ms = reinterpret_cast<Hnsw*>(h)->last_kernel_ms;
  return COLTT_OK;
}
COLTT_API int coltt_b200_hnsw_last_stats(coltt_hnsw* h, uint64_t* dist_evals, uint64_t* expansions) {
  if (!h || !dist_evals || !expansions) return fail(COLTT_ERR_INVALID, "null argument");
  *dist_evals = reinterpret_cast<Hnsw*>(h)->last_evals;
  *expansions = reinterpret_cast<Hnsw*>(h)->last_exp;
  return COLTT_OK;
}
}
