// comm.cu — the multi-GPU exchange step of a sharded collection, behind the C-ABI.
//
// Reference analogue: a vectorspace scans its 16 in-process map shards into shard-local queues and re-Adds them into
// one queue (edge/none_vectorstore.go:152-178); rows -> shard by ShardVertex (pkg/sharding/shard.go:34-41).  Here a
// shard is one GPU's coltt_store (or coltt_hnsw sub-graph) and the re-merge is ONE ncclAllGather of the
// per-shard top-k lists over NVLink followed by the K5 merge kernel (topk_merge.cu) on every rank.  There is no other
// collective in the path.
//
// A coltt_comm is one rank: a device, its NCCL communicator, a stream and persistent send / receive / pinned staging
// buffers (no allocation per call).  The local search writes its hits and counts straight into the send buffer (one
// message per rank: coltt_hit[nq][k] then int32 counts[nq]), so nothing is repacked between the search and the
// all-gather.  Two ways to build the ranks:
//   * one process per GPU (bench.py under torchrun): coltt_b200_comm_unique_id on rank 0, the 128-byte blob travels
//     through the host's own channel, coltt_b200_comm_init_rank everywhere;
//   * one process driving all GPUs (the Go host): coltt_b200_init(device_ids, n) = ncclCommInitAll, and
//     either n threads (goroutines) calling coltt_b200_sharded_search with their rank's handle, or one call of
//     coltt_b200_sharded_search_all, which runs the ranks on n library threads.
// The exchange itself does not have to be an NCCL call.  When every pair of ranks can map the other's memory (NVLink /
// NVSwitch peers: cudaIpc handles between processes, cudaDeviceEnablePeerAccess inside one) the all-gather is FUSED INTO
// THE MERGE: each rank's local search writes its hits into its own exchange buffer and raises a flag (system-scope
// release); a one-CTA kernel on every rank waits for the peers' flags; then K5 reads the peers' lists straight out of
// their memory over NVLink while it merges — no gather kernel, no receive buffer, no second pass.  Exchange buffers are
// double-buffered by the parity of a per-communicator sequence number (a rank can only reach step s+2 after every peer
// has finished merging step s).  NCCL then only bootstraps (it carries the IPC handles once) and stays as the fallback
// (COLTT_P2P=0, more than 8 ranks, no peer access).  The wait has a timeout: a peer that never arrives is an error code,
// not a hung GPU.
// NCCL is loaded at run time (dlopen of libnccl.so.2, preferring the copy already mapped into the process), so
// libcoltt_b200.so has no link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <mutex>
#include <thread>
#include <vector>

#include "hnsw.h"
#include "kernels.cuh"
#include "store.h"

namespace coltt {

int hnsw_search_keep_device(Hnsw* h, const float* queries, size_t nq, int k, int ef, const Hit** d_hits, const int** d_counts,
                            cudaStream_t* st);   // hnsw.cu
int hnsw_pq_search(Hnsw* h, const float* queries, size_t nq, int k, int ef_in, int rerank, uint64_t* out_ids, float* out_scores,
                   int32_t* out_counts);         // pq.cu

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
  std::string why;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the host process already uses, if any
    if (!h) { const char* e = getenv("COLTT_NCCL_LIB"); if (e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?"); return; }
    bool all = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { all = false; api.why = std::string("NCCL symbol missing: ") + name; } return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = all;
  });
  return api;
}

#define COLTT_NCCL(expr)                                                                                              \
  do {                                                                                                                \
    ncclResult_t _r = (expr);                                                                                         \
    if (_r != ncclSuccess) return ::coltt::fail(COLTT_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
  } while (0)

struct Comm {
  int device = 0, rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  std::mutex mu;                       // one collective at a time per rank
  DeviceBuf send, recv, q_in, out, counts;
  PinnedBuf h_q, h_out;
  // peer-memory exchange (see the header comment)
  int p2p = 0;                         // 0 = not set up yet, 1 = active, -1 = unavailable: NCCL path
  uint8_t* xbuf = nullptr;             // [2][xcap] this rank's messages of even / odd steps
  size_t xcap = 0;
  uint32_t* flags = nullptr;           // [0..1] sequence number last published per slot
  uint32_t* h_marker = nullptr;        // page-locked, device-mapped: the step at which the wait kernel gave up on a peer (0 = never).
                                       // The host reads it without a copy or a synchronisation; once set the communicator is broken.
  std::vector<void*> peer_x, peer_f, ipc_opened;
  DeviceBuf d_bases[2], d_peer_flags;  // device tables: list base addresses per slot; peers' flag words
  uint32_t seq = 0;
  ~Comm() {
    cudaSetDevice(device);
    for (void* q : ipc_opened) cudaIpcCloseMemHandle(q);
    if (xbuf) cudaFree(xbuf);
    if (flags) cudaFree(flags);
    if (h_marker) cudaFreeHost(h_marker);
    if (comm && nccl().ok) nccl().CommDestroy(comm);
    if (stream) cudaStreamDestroy(stream);
  }
};

static std::mutex g_all_mu;
static std::vector<Comm*> g_all;       // ranks created by coltt_b200_init (destroyed by coltt_b200_shutdown)

static size_t msg_bytes(size_t nq, int k) { return (nq * (size_t)k * sizeof(Hit) + nq * 4 + 15) / 16 * 16; }

static int make_comm(int device, int rank, int world, ncclComm_t c, Comm** out) {
  std::unique_ptr<Comm> cm(new Comm());
  cm->device = device; cm->rank = rank; cm->world = world; cm->comm = c;
  COLTT_CUDA(cudaSetDevice(device));
  COLTT_CUDA(cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking));
  *out = cm.release();
  return COLTT_OK;
}

static int ensure_bufs(Comm& cm, size_t nq, int k) {
  const size_t mb = msg_bytes(nq, k);
  int rc;
  if ((rc = cm.send.ensure(mb))) return rc;
  if (cm.world > 1 && (rc = cm.recv.ensure(mb * cm.world))) return rc;
  return COLTT_OK;
}
// ---- peer-memory exchange ------------------------------------------------------------------------------------------
__global__ void xchg_signal_kernel(uint32_t* flag, uint32_t seq) {
  // everything earlier in the stream (the local search's writes into the exchange buffer) is complete; publish it to the peers
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
}
__global__ void xchg_wait_kernel(const unsigned long long* peer_flags, int world, uint32_t slot, uint32_t seq, uint32_t* timeout_marker,
                                 long long timeout_cycles) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const uint32_t* f = reinterpret_cast<const uint32_t*>(peer_flags[r]) + slot;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int32_t)(v - seq) >= 0) break;                  // sequence numbers only grow
    if (clock64() - t0 > timeout_cycles) {               // the marker lives in mapped host memory: plain store + system fence
      *reinterpret_cast<volatile uint32_t*>(timeout_marker) = seq ? seq : 1u;
      __threadfence_system();
      break;
    }
    __nanosleep(200);
  }
}

struct XchgInfo {                      // what the ranks tell each other once, through NCCL
  unsigned long long pid, x_ptr, f_ptr, xcap;
  cudaIpcMemHandle_t hx, hf;
  unsigned long long ok;
};

// Collective: (re)creates the exchange buffers for messages of `need` bytes and maps every peer's.  All ranks call it at the
// same step (the message size is a function of nq and k, which the collective contract makes equal everywhere).
static int p2p_setup(Comm& cm, size_t need) {
  static const bool enabled = [] { const char* e = getenv("COLTT_P2P"); return !e || atoi(e) != 0; }();
  if (!enabled || cm.world < 2 || cm.world > 8) { cm.p2p = -1; return COLTT_OK; }
  cudaStream_t st = cm.stream;
  int rc;
  // Barrier first: a peer may still be merging the previous step out of THIS rank's buffer.  Every rank synchronises its own
  // stream and then joins a tiny all-gather, which completes only when all of them have — after that nobody reads anybody.
  COLTT_CUDA(cudaDeviceSynchronize());
  if ((rc = cm.send.ensure(sizeof(XchgInfo))) || (rc = cm.recv.ensure(sizeof(XchgInfo) * cm.world))) return rc;
  COLTT_NCCL(nccl().AllGather(cm.send.p, cm.recv.p, 8, ncclChar, cm.comm, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  for (void* q : cm.ipc_opened) cudaIpcCloseMemHandle(q);
  cm.ipc_opened.clear(); cm.peer_x.assign(cm.world, nullptr); cm.peer_f.assign(cm.world, nullptr);
  if (cm.xbuf) { cudaFree(cm.xbuf); cm.xbuf = nullptr; }
  if (!cm.flags) { COLTT_CUDA(cudaMalloc(&cm.flags, 64)); COLTT_CUDA(cudaMemset(cm.flags, 0, 64)); cm.seq = 0; }
  if (!cm.h_marker) {
    COLTT_CUDA(cudaHostAlloc((void**)&cm.h_marker, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *cm.h_marker = 0;
  }
  const size_t cap = std::max<size_t>((need + 255) / 256 * 256, 1u << 20);
  XchgInfo mine{};
  mine.ok = 1;
  if (cudaMalloc(&cm.xbuf, 2 * cap) != cudaSuccess) { cudaGetLastError(); cm.xbuf = nullptr; mine.ok = 0; }
  cm.xcap = cap;
  mine.pid = (unsigned long long)getpid(); mine.x_ptr = (unsigned long long)cm.xbuf; mine.f_ptr = (unsigned long long)cm.flags; mine.xcap = cap;
  if (mine.ok && (cudaIpcGetMemHandle(&mine.hx, cm.xbuf) != cudaSuccess || cudaIpcGetMemHandle(&mine.hf, cm.flags) != cudaSuccess)) {
    cudaGetLastError();
    mine.ok = 0;
  }
  // one NCCL all-gather of the descriptors (the sequence number simply continues: it is the same on every rank)
  const size_t ib = sizeof(XchgInfo);
  std::vector<XchgInfo> all(cm.world);
  COLTT_CUDA(cudaMemcpyAsync(cm.send.p, &mine, ib, cudaMemcpyHostToDevice, st));
  COLTT_NCCL(nccl().AllGather(cm.send.p, cm.recv.p, ib, ncclChar, cm.comm, st));
  COLTT_CUDA(cudaMemcpyAsync(all.data(), cm.recv.p, ib * cm.world, cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  bool ok = true;
  for (int r = 0; r < cm.world; r++) ok = ok && all[r].ok && all[r].xcap == cap;
  if (ok) {
    for (int r = 0; r < cm.world && ok; r++) {
      if (r == cm.rank) { cm.peer_x[r] = cm.xbuf; cm.peer_f[r] = cm.flags; continue; }
      if (all[r].pid == mine.pid) {           // same process (coltt_b200_init): plain peer access
        int peer_dev = -1;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, (void*)all[r].x_ptr) == cudaSuccess) peer_dev = at.device;
        int can = 0;
        if (peer_dev < 0 || cudaDeviceCanAccessPeer(&can, cm.device, peer_dev) != cudaSuccess || !can) { ok = false; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = false; break; }
        cudaGetLastError();
        cm.peer_x[r] = (void*)all[r].x_ptr; cm.peer_f[r] = (void*)all[r].f_ptr;
      } else {
        void *px = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&px, all[r].hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
        cm.ipc_opened.push_back(px);
        if (cudaIpcOpenMemHandle(&pf, all[r].hf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
        cm.ipc_opened.push_back(pf);
        cm.peer_x[r] = px; cm.peer_f[r] = pf;
      }
    }
    cudaGetLastError();
  }
  // consensus: the peer path is used only if EVERY rank mapped every peer (a mixed decision would deadlock)
  unsigned long long vote = ok ? 1 : 0;
  std::vector<unsigned long long> votes(cm.world);
  COLTT_CUDA(cudaMemcpyAsync(cm.send.p, &vote, 8, cudaMemcpyHostToDevice, st));
  COLTT_NCCL(nccl().AllGather(cm.send.p, cm.recv.p, 8, ncclChar, cm.comm, st));
  COLTT_CUDA(cudaMemcpyAsync(votes.data(), cm.recv.p, 8 * cm.world, cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < cm.world; r++) ok = ok && votes[r] == 1;
  if (!ok) { cm.p2p = -1; return COLTT_OK; }
  // device tables: list bases per slot, peers' flag words
  std::vector<unsigned long long> b0(cm.world), b1(cm.world), pf(cm.world);
  for (int r = 0; r < cm.world; r++) { b0[r] = (unsigned long long)cm.peer_x[r]; b1[r] = b0[r] + cap; pf[r] = (unsigned long long)cm.peer_f[r]; }
  if ((rc = cm.d_bases[0].ensure(8 * cm.world)) || (rc = cm.d_bases[1].ensure(8 * cm.world)) || (rc = cm.d_peer_flags.ensure(8 * cm.world))) return rc;
  COLTT_CUDA(cudaMemcpyAsync(cm.d_bases[0].p, b0.data(), 8 * cm.world, cudaMemcpyHostToDevice, st));
  COLTT_CUDA(cudaMemcpyAsync(cm.d_bases[1].p, b1.data(), 8 * cm.world, cudaMemcpyHostToDevice, st));
  COLTT_CUDA(cudaMemcpyAsync(cm.d_peer_flags.p, pf.data(), 8 * cm.world, cudaMemcpyHostToDevice, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  cm.p2p = 1;
  return COLTT_OK;
}

// Where this step's local search must write its hits (+ counts): the rank's exchange slot, or the NCCL send buffer.
static int exchange_begin(Comm& cm, size_t nq, int k, uint8_t** msg) {
  const size_t mb = msg_bytes(nq, k);
  if (cm.world > 1 && nccl().ok && (cm.p2p == 0 || (cm.p2p == 1 && mb > cm.xcap))) {
    int rc = p2p_setup(cm, mb);
    if (rc) return rc;
  }
  if (cm.p2p == 1) {
    cm.seq++;
    *msg = cm.xbuf + (size_t)(cm.seq & 1u) * cm.xcap;
    return COLTT_OK;
  }
  int rc = ensure_bufs(cm, nq, k);
  if (rc) return rc;
  *msg = (uint8_t*)cm.send.p;
  return COLTT_OK;
}

// The exchange: this rank's message is already in cm.send; all-gather, then merge the `world` lists per query.
static int exchange_and_merge(Comm& cm, size_t nq, int k, int nearest, Hit* d_out, int* d_counts, cudaStream_t st) {
  const size_t mb = msg_bytes(nq, k);
  if (cm.p2p == 1) {
    // fused exchange: publish, wait for the peers' flags (one small CTA), merge straight out of the peers' memory
    const uint32_t slot = cm.seq & 1u;
    xchg_signal_kernel<<<1, 1, 0, st>>>(cm.flags + slot, cm.seq);
    static const long long timeout_cycles = [] { const char* e = getenv("COLTT_P2P_TIMEOUT_MS"); return (long long)(e ? atoi(e) : 20000) * 1500000ll; }();
    xchg_wait_kernel<<<1, 32, 0, st>>>((const unsigned long long*)cm.d_peer_flags.p, cm.world, slot, cm.seq, cm.h_marker, timeout_cycles);
    count_launch(2);
    COLTT_CUDA(cudaGetLastError());
    MergeParams mp{};
    mp.list_bases = (const unsigned long long*)cm.d_bases[slot].p; mp.counts_off = nq * (size_t)k * sizeof(Hit);
    mp.lists = (const Hit*)cm.xbuf; mp.counts = (const int*)cm.xbuf;     // unused with list_bases
    mp.n_lists = cm.world; mp.nq = (uint32_t)nq; mp.k_in = (uint32_t)k; mp.k = (uint32_t)k; mp.nearest = nearest; mp.in_best_first = 0;
    mp.out = d_out; mp.out_counts = d_counts;
    return launch_merge_topk(mp, st);
  }
  if (cm.world > 1) COLTT_NCCL(nccl().AllGather(cm.send.p, cm.recv.p, mb, ncclChar, cm.comm, st));
  MergeParams mp{};
  const uint8_t* base = (const uint8_t*)(cm.world > 1 ? cm.recv.p : cm.send.p);
  mp.lists = (const Hit*)base; mp.counts = (const int*)(base + nq * (size_t)k * sizeof(Hit)); mp.n_lists = cm.world;
  mp.nq = (uint32_t)nq; mp.k_in = (uint32_t)k; mp.k = (uint32_t)k; mp.nearest = nearest; mp.in_best_first = 0;
  mp.list_stride_hits = mb / 16; mp.count_stride = mb / 4;
  mp.out = d_out; mp.out_counts = d_counts;
  return launch_merge_topk(mp, st);
}


// Did the wait kernel give up on a peer?  Checked after every synchronised host call, and at the start of every call for
// the enqueue-only entry points (whose failure can only surface at the next call).  Sticky: the ranks' sequence numbers no
// longer agree after a timeout, so the communicator stays failed until it is destroyed.
static int check_exchange_timeout(Comm& cm) {
  if (cm.p2p != 1 || !cm.h_marker) return COLTT_OK;
  const uint32_t marker = *reinterpret_cast<volatile uint32_t*>(cm.h_marker);
  if (marker) return fail(COLTT_ERR_CUDA, "sharded search: a peer rank did not publish its results within the exchange timeout (step " + std::to_string(marker) + ")");
  return COLTT_OK;
}

// Enqueue-only: local search into the send buffer, all-gather, merge into d_out / d_counts, all on `st`.
static int sharded_search_enqueue(Comm& cm, Store* shard, const void* d_queries, size_t nq, int k, int select_mode, int math_mode,
                                  Hit* d_out, int* d_counts, cudaStream_t st) {
  if (select_mode != COLTT_SELECT_COMPAT && select_mode != COLTT_SELECT_NEAREST) return fail(COLTT_ERR_INVALID, "bad select mode");
  if (shard->device != cm.device) return fail(COLTT_ERR_INVALID, "shard and communicator live on different devices");
  uint8_t* msg = nullptr;
  int rc = check_exchange_timeout(cm);
  if (rc) return rc;
  if ((rc = exchange_begin(cm, nq, k, &msg))) return rc;
  Hit* s_hits = (Hit*)msg;
  int* s_cnt = (int*)(msg + nq * (size_t)k * sizeof(Hit));
  rc = shard->search_dev(d_queries, nq, k, select_mode, math_mode, s_hits, s_cnt, st);   // caller stream: enqueue only
  if (rc) return rc;
  return exchange_and_merge(cm, nq, k, select_mode == COLTT_SELECT_NEAREST, d_out, d_counts, st);
}

static void unpack(const Hit* h, const int* c, size_t nq, int k, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  for (size_t q = 0; q < nq; q++) {
    out_counts[q] = c[q];
    for (int i = 0; i < c[q]; i++) {
      out_ids[q * (size_t)k + i] = h[q * (size_t)k + i].id;
      out_scores[q * (size_t)k + i] = h[q * (size_t)k + i].score;
    }
  }
}

static int sharded_search_host(Comm& cm, Store* shard, const float* queries, size_t nq, int k, int select_mode, int math_mode,
                               uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || k <= 0) return fail(COLTT_ERR_INVALID, "bad argument");
  std::lock_guard<std::mutex> g(cm.mu);
  COLTT_CUDA(cudaSetDevice(cm.device));
  const size_t qb = nq * (size_t)shard->dim * 4, hb = nq * (size_t)k * sizeof(Hit);
  int rc;
  if ((rc = cm.q_in.ensure(qb)) || (rc = cm.out.ensure(hb)) || (rc = cm.counts.ensure(nq * 4)) || (rc = cm.h_q.ensure(qb)) ||
      (rc = cm.h_out.ensure(hb + nq * 4)))
    return rc;
  const bool direct = host_ptr_is_pinned(queries);     // a page-locked caller buffer is the DMA source itself (store.cu)
  if (!direct) std::memcpy(cm.h_q.p, queries, qb);
  COLTT_CUDA(cudaMemcpyAsync(cm.q_in.p, direct ? (const void*)queries : (const void*)cm.h_q.p, qb, cudaMemcpyHostToDevice, cm.stream));
  rc = sharded_search_enqueue(cm, shard, cm.q_in.p, nq, k, select_mode, math_mode, (Hit*)cm.out.p, (int*)cm.counts.p, cm.stream);
  if (rc) return rc;
  const bool want = out_ids && out_scores && out_counts;     // ranks other than the caller's front rank may pass NULL outputs
  if (want) {
    COLTT_CUDA(cudaMemcpyAsync(cm.h_out.p, cm.out.p, hb, cudaMemcpyDeviceToHost, cm.stream));
    COLTT_CUDA(cudaMemcpyAsync((uint8_t*)cm.h_out.p + hb, cm.counts.p, nq * 4, cudaMemcpyDeviceToHost, cm.stream));
  }
  COLTT_CUDA(cudaStreamSynchronize(cm.stream));
  if ((rc = check_exchange_timeout(cm))) return rc;
  if (want) unpack((const Hit*)cm.h_out.p, (const int*)((uint8_t*)cm.h_out.p + hb), nq, k, out_ids, out_scores, out_counts);
  return COLTT_OK;
}

// HNSW shards: one independent sub-graph per GPU over its row shard; the same all-gather + merge (nearest first).
// pq: 0 = fp32 walk (hnsw.cu), 1 = PQ walk + exact re-rank, 2 = PQ walk, quantized scores (pq.cu)
static int sharded_hnsw_host(Comm& cm, Hnsw* h, const float* queries, size_t nq, int k, int ef, int pq, uint64_t* out_ids, float* out_scores,
                             int32_t* out_counts) {
  if (nq == 0) return COLTT_OK;
  if (!queries || k <= 0) return fail(COLTT_ERR_INVALID, "bad argument");
  if (h->device != cm.device) return fail(COLTT_ERR_INVALID, "sub-graph and communicator live on different devices");
  std::lock_guard<std::mutex> g(cm.mu);
  COLTT_CUDA(cudaSetDevice(cm.device));
  const size_t hb = nq * (size_t)k * sizeof(Hit);
  int rc;
  uint8_t* msg = nullptr;
  if ((rc = exchange_begin(cm, nq, k, &msg)) || (rc = cm.out.ensure(hb)) || (rc = cm.counts.ensure(nq * 4)) || (rc = cm.h_out.ensure(hb + nq * 4))) return rc;
  const Hit* d_hits = nullptr; const int* d_cnt = nullptr; cudaStream_t hst = nullptr;
  if (pq) {
    std::vector<uint64_t> ti(nq * (size_t)k);
    std::vector<float> ts(nq * (size_t)k);
    std::vector<int32_t> tc(nq);
    rc = hnsw_pq_search(h, queries, nq, k, ef, pq == 1, ti.data(), ts.data(), tc.data());   // the hits also stay in the handle's scratch
    if (rc) return rc;
    if (h->n == 0) return fail(COLTT_ERR_UNSUPPORTED, "empty sub-graph in a sharded PQ search");
    d_hits = (const Hit*)h->out.p; d_cnt = (const int*)h->counts.p;
  } else {
    rc = hnsw_search_keep_device(h, queries, nq, k, ef, &d_hits, &d_cnt, &hst);     // returns after the walk finished
    if (rc) return rc;
  }
  COLTT_CUDA(cudaMemcpyAsync(msg, d_hits, hb, cudaMemcpyDeviceToDevice, cm.stream));
  COLTT_CUDA(cudaMemcpyAsync(msg + hb, d_cnt, nq * 4, cudaMemcpyDeviceToDevice, cm.stream));
  rc = exchange_and_merge(cm, nq, k, 1, (Hit*)cm.out.p, (int*)cm.counts.p, cm.stream);
  if (rc) return rc;
  const bool want = out_ids && out_scores && out_counts;
  if (want) {
    COLTT_CUDA(cudaMemcpyAsync(cm.h_out.p, cm.out.p, hb, cudaMemcpyDeviceToHost, cm.stream));
    COLTT_CUDA(cudaMemcpyAsync((uint8_t*)cm.h_out.p + hb, cm.counts.p, nq * 4, cudaMemcpyDeviceToHost, cm.stream));
  }
  COLTT_CUDA(cudaStreamSynchronize(cm.stream));
  if ((rc = check_exchange_timeout(cm))) return rc;
  if (want) unpack((const Hit*)cm.h_out.p, (const int*)((uint8_t*)cm.h_out.p + hb), nq, k, out_ids, out_scores, out_counts);
  return COLTT_OK;
}

}  // namespace coltt

using coltt::Comm;
using coltt::fail;
using coltt::nccl;

extern "C" {

COLTT_API int coltt_b200_comm_unique_id(void* out, size_t len) {
  if (!out || len < sizeof(ncclUniqueId)) return fail(COLTT_ERR_INVALID, "unique id buffer must hold 128 bytes");
  if (!nccl().ok) return fail(COLTT_ERR_UNSUPPORTED, nccl().why);
  ncclUniqueId id;
  COLTT_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out, &id, sizeof(id));
  return COLTT_OK;
}

COLTT_API int coltt_b200_comm_init_rank(const void* unique_id, int rank, int world, int device, coltt_comm** out) {
  if (!unique_id || !out || world < 1 || rank < 0 || rank >= world) return fail(COLTT_ERR_INVALID, "bad communicator arguments");
  int rc = coltt::require_device(device);
  if (rc) return rc;
  if (!nccl().ok) return fail(COLTT_ERR_UNSUPPORTED, nccl().why);
  COLTT_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  ncclComm_t c = nullptr;
  COLTT_NCCL(nccl().CommInitRank(&c, world, id, rank));
  Comm* cm = nullptr;
  rc = coltt::make_comm(device, rank, world, c, &cm);
  if (rc) { nccl().CommDestroy(c); return rc; }
  *out = reinterpret_cast<coltt_comm*>(cm);
  return COLTT_OK;
}

COLTT_API int coltt_b200_init(const int* device_ids, int n_dev, coltt_comm** out_comms) {
  if (!device_ids || !out_comms || n_dev < 1 || n_dev > 64) return fail(COLTT_ERR_INVALID, "bad device list");
  for (int i = 0; i < n_dev; i++) { int rc = coltt::require_device(device_ids[i]); if (rc) return rc; }
  std::vector<ncclComm_t> cs(n_dev, nullptr);
  if (n_dev > 1) {
    if (!nccl().ok) return fail(COLTT_ERR_UNSUPPORTED, nccl().why);
    COLTT_NCCL(nccl().CommInitAll(cs.data(), n_dev, device_ids));
  }
  std::lock_guard<std::mutex> g(coltt::g_all_mu);
  for (int i = 0; i < n_dev; i++) {
    Comm* cm = nullptr;
    int rc = coltt::make_comm(device_ids[i], i, n_dev, cs[i], &cm);
    if (rc) return rc;
    coltt::g_all.push_back(cm);
    out_comms[i] = reinterpret_cast<coltt_comm*>(cm);
  }
  return COLTT_OK;
}

COLTT_API void coltt_b200_comm_destroy(coltt_comm* c) {
  Comm* cm = reinterpret_cast<Comm*>(c);
  if (!cm) return;
  {
    std::lock_guard<std::mutex> g(coltt::g_all_mu);
    for (auto it = coltt::g_all.begin(); it != coltt::g_all.end(); ++it)
      if (*it == cm) { coltt::g_all.erase(it); break; }
  }
  delete cm;
}

COLTT_API void coltt_b200_shutdown(void) {
  std::vector<Comm*> all;
  {
    std::lock_guard<std::mutex> g(coltt::g_all_mu);
    all.swap(coltt::g_all);
  }
  // peers read each other's exchange buffers while they merge: every device is idle before any rank's buffers are freed
  for (Comm* cm : all)
    if (cudaSetDevice(cm->device) == cudaSuccess) cudaDeviceSynchronize();
  cudaGetLastError();
  for (Comm* cm : all) delete cm;
}

COLTT_API int coltt_b200_comm_info(coltt_comm* c, int* rank, int* world, int* device) {
  Comm* cm = reinterpret_cast<Comm*>(c);
  if (!cm) return fail(COLTT_ERR_INVALID, "null communicator");
  if (rank) *rank = cm->rank;
  if (world) *world = cm->world;
  if (device) *device = cm->device;
  return COLTT_OK;
}

COLTT_API int coltt_b200_comm_exchange_mode(coltt_comm* c) {
  Comm* cm = reinterpret_cast<Comm*>(c);
  if (!cm) return fail(COLTT_ERR_INVALID, "null communicator");
  if (cm->world < 2) return COLTT_EXCHANGE_NCCL;
  return cm->p2p == 1 ? COLTT_EXCHANGE_PEER : cm->p2p == 0 ? COLTT_EXCHANGE_UNDECIDED : COLTT_EXCHANGE_NCCL;
}

COLTT_API int coltt_b200_sharded_search(coltt_comm* c, coltt_store* shard, const float* queries, size_t nq, int k, int select_mode,
                                        int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (!c || !shard) return fail(COLTT_ERR_INVALID, "null handle");
  return coltt::sharded_search_host(*reinterpret_cast<Comm*>(c), reinterpret_cast<coltt::Store*>(shard), queries, nq, k, select_mode, math_mode,
                                    out_ids, out_scores, out_counts);
}

COLTT_API int coltt_b200_sharded_search_dev(coltt_comm* c, coltt_store* shard, const void* d_queries, size_t nq, int k, int select_mode,
                                            int math_mode, void* d_out, void* d_counts, void* stream) {
  if (!c || !shard || !d_queries || !d_out || !d_counts) return fail(COLTT_ERR_INVALID, "null argument");
  if (nq == 0) return COLTT_OK;
  if (k <= 0) return fail(COLTT_ERR_INVALID, "top-k must be positive");
  Comm& cm = *reinterpret_cast<Comm*>(c);
  std::lock_guard<std::mutex> g(cm.mu);
  COLTT_CUDA(cudaSetDevice(cm.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : cm.stream;
  int rc = coltt::sharded_search_enqueue(cm, reinterpret_cast<coltt::Store*>(shard), d_queries, nq, k, select_mode, math_mode, (coltt::Hit*)d_out,
                                         (int*)d_counts, st);
  if (rc) return rc;
  if (!stream) {
    COLTT_CUDA(cudaStreamSynchronize(st));
    return coltt::check_exchange_timeout(cm);
  }
  return COLTT_OK;
}

COLTT_API int coltt_b200_sharded_search_all(coltt_comm* const* comms, coltt_store* const* shards, int n, const float* queries, size_t nq, int k,
                                            int select_mode, int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (!comms || !shards || n < 1) return fail(COLTT_ERR_INVALID, "bad argument");
  std::vector<int> rcs(n, COLTT_OK);
  std::vector<std::string> msgs(n);
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++)
    th.emplace_back([&, i] {
      rcs[i] = coltt_b200_sharded_search(comms[i], shards[i], queries, nq, k, select_mode, math_mode, nullptr, nullptr, nullptr);
      if (rcs[i]) msgs[i] = coltt::last_error_cstr();
    });
  rcs[0] = coltt_b200_sharded_search(comms[0], shards[0], queries, nq, k, select_mode, math_mode, out_ids, out_scores, out_counts);
  for (auto& t : th) t.join();
  for (int i = 1; i < n; i++)
    if (rcs[i]) return fail(rcs[i], "rank " + std::to_string(i) + ": " + msgs[i]);
  return rcs[0];
}

COLTT_API int coltt_b200_sharded_hnsw_search(coltt_comm* c, coltt_hnsw* sub, const float* queries, size_t nq, int k, int ef, uint64_t* out_ids,
                                             float* out_scores, int32_t* out_counts) {
  if (!c || !sub) return fail(COLTT_ERR_INVALID, "null handle");
  return coltt::sharded_hnsw_host(*reinterpret_cast<Comm*>(c), reinterpret_cast<coltt::Hnsw*>(sub), queries, nq, k, ef, 0, out_ids, out_scores,
                                  out_counts);
}

COLTT_API int coltt_b200_sharded_hnsw_pq_search(coltt_comm* c, coltt_hnsw* sub, const float* queries, size_t nq, int k, int ef, int rerank,
                                                uint64_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (!c || !sub) return fail(COLTT_ERR_INVALID, "null handle");
  return coltt::sharded_hnsw_host(*reinterpret_cast<Comm*>(c), reinterpret_cast<coltt::Hnsw*>(sub), queries, nq, k, ef, rerank ? 1 : 2, out_ids,
                                  out_scores, out_counts);
}

}  // extern "C"
