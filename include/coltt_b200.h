/* coltt_b200.h — C-ABI of libcoltt_b200.so: the B200-native ANN search path that drops in
 * behind coltt's edge Vectorstore/Quantization and core/vectorindex HNSW interfaces.
 *
 * The reference (sjy-dv/coltt) is pure Go with no FFI boundary; the seams this library
 * replaces are Go interfaces.  Each entry point below names the reference method(s) a cgo
 * shim would forward to it (INTEGRATION.md shows the shim).  Plain pointers and sizes only;
 * no torch/CUDA types in signatures (device pointers travel as void*).
 *
 * Conventions: every function returns COLTT_OK (0) or a negative coltt_status and never
 * aborts or throws across the boundary; coltt_b200_last_error() returns a thread-local
 * message for the last failure on the calling thread.  *_search calls on one handle may run
 * concurrently from many threads (the reference searches under per-shard RLock,
 * edge/none_vectorstore.go:137-146); upsert/remove/import are serialized per handle
 * (reference: one shard Lock, none_vectorstore.go:99-101).  Caller-owned buffers are
 * copied in and never retained (cgo pointer rules).  There is NO CPU fallback: if no
 * sm_100 device is present every call that needs one fails with COLTT_ERR_NO_DEVICE.
 */
#ifndef COLTT_B200_H
#define COLTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define COLTT_API __attribute__((visibility("default")))
#else
#define COLTT_API
#endif

typedef enum coltt_status {
  COLTT_OK = 0,
  COLTT_ERR_INVALID = -1,     /* bad argument */
  COLTT_ERR_CUDA = -2,        /* CUDA runtime/driver error (message has the detail) */
  COLTT_ERR_NOMEM = -3,
  COLTT_ERR_NOT_FOUND = -4,
  COLTT_ERR_DIM = -5,         /* "Dim Length UnmatchdError" (none_vectorstore.go:86-88) */
  COLTT_ERR_UNSUPPORTED = -6,
  COLTT_ERR_FORMAT = -7,      /* malformed SaveVertex / Commit blob */
  COLTT_ERR_NO_DEVICE = -8    /* no sm_100 GPU: there is no CPU fallback */
} coltt_status;

/* edgepb.Distance (idl/proto/v4/edge.proto:69-72) */
typedef enum coltt_metric { COLTT_COSINE = 0, COLTT_EUCLIDEAN = 1 } coltt_metric;

/* edgepb.Quantization (edge.proto:75-80).  BF16 is stored and decoded as IEEE binary16
 * because the reference's "bf16" codec is (pkg/compresshelper/bf16.go:233-317).
 * F8 is the reference's literal (broken) 8-bit code (float8.go:233-313).
 * F8_E4M3 is a builder extension with no reference arithmetic (the reference's f8 is unusable): rows are
 * OCP E4M3 codes (round-to-nearest-even, saturating) plus one power-of-two scale per row,
 *   s = 2^clamp(floor(log2(max|v|)) - 7, -40, 40),  code_i = e4m3(v_i / s),  value_i = s * decode(code_i),
 * the query is lowered the same way (as f8_vectorstore.go:136-139 lowers it) and the distance is the reference's
 * arithmetic over the dequantized values (f8_quantization.go:33-43 -> pkg/distance).  SaveVertex/LoadVertex have no
 * layout for it (export/import return COLTT_ERR_UNSUPPORTED). */
typedef enum coltt_quant {
  COLTT_QUANT_NONE = 0,
  COLTT_QUANT_F16 = 1,
  COLTT_QUANT_F8 = 2,
  COLTT_QUANT_BF16 = 3,
  COLTT_QUANT_F8_E4M3 = 16
} coltt_quant;

/* Which K rows a FLAT search keeps.
 * COLTT_SELECT_COMPAT: the reference's literal behaviour — edge.PriorityQueue over a MIN heap
 *   pops the minimum distance when over capacity, i.e. keeps the K LARGEST distances, returned
 *   ascending (edge/priority_queue.go:39-69).
 * COLTT_SELECT_NEAREST: keeps the K smallest distances, ascending (what HNSW and users expect).
 * Both are windows of ONE total order T = (score ascending, NaN after every number, then id
 * ascending): NEAREST returns the first K entries of T, COMPAT the last K, both listed in T order.
 * (The reference is nondeterministic among equal scores — Go map iteration, F6 — T pins it.) */
typedef enum coltt_select { COLTT_SELECT_COMPAT = 0, COLTT_SELECT_NEAREST = 1 } coltt_select;

/* COLTT_MATH_EXACT: CUDA-core kernel that reproduces the reference AVX evaluation order
 *   (pkg/distance/simd/cpp/avx.cpp:3-8,15-32,51-75) — scores bit-identical to the Go path.
 * COLTT_MATH_FAST: tcgen05 tensor-core filter over the whole shard, then the survivors are
 *   re-scored by the EXACT arithmetic; a certified margin makes the returned ids/scores equal
 *   to EXACT, and queries whose margin cannot be certified are re-run EXACT automatically. */
typedef enum coltt_math { COLTT_MATH_EXACT = 0, COLTT_MATH_FAST = 1 } coltt_math;

typedef struct coltt_store coltt_store; /* one edge collection's vectors on one GPU */
typedef struct coltt_hnsw coltt_hnsw;   /* one core/vectorindex.Hnsw on one GPU */

typedef struct coltt_store_cfg {
  uint32_t dim;            /* Collection.dim (edge.proto:34) */
  int32_t metric;          /* coltt_metric */
  int32_t quant;           /* coltt_quant */
  int32_t device;          /* CUDA device ordinal */
  uint64_t capacity_hint;  /* rows to reserve up front (0 = grow on demand) */
} coltt_store_cfg;

/* ---- library ------------------------------------------------------------------------- */
COLTT_API const char* coltt_b200_last_error(void);
COLTT_API const char* coltt_b200_version(void);
/* Number of visible sm_100 devices (0 when none; never an error). */
COLTT_API int coltt_b200_device_count(void);

/* ---- edge FLAT store: replaces the `vectorspace` implementations ------------------------
 * newNoneVectorstore / newF16Vectorstore / newBF16Vectorstore / newF8Vectorstore
 * (edge/vectorstore.go:62-85). */
COLTT_API int coltt_b200_store_create(const coltt_store_cfg* cfg, coltt_store** out);
COLTT_API void coltt_b200_store_destroy(coltt_store* s);
/* vectorspace.LoadSize / Dim (edge/vectorstore.go:44-46) */
COLTT_API int coltt_b200_store_size(coltt_store* s, uint64_t* n_rows);
COLTT_API int coltt_b200_store_dim(coltt_store* s, uint32_t* dim);

/* vectorspace.ChangedVertex (edge/none_vectorstore.go:66-103, bf16_vectorstore.go:66-106):
 * n rows of `dim` un-normalized fp32; normalized (cosine) and Lower()ed on the GPU with the
 * reference's arithmetic; an existing id is overwritten in place.  Metadata / inverted index
 * stay on the Go side.  `vecs` is a host pointer. */
COLTT_API int coltt_b200_store_upsert(coltt_store* s, const uint64_t* ids, const float* vecs, size_t n);
/* vectorspace.RemoveVertex after the Go side resolved dropFilter to ids
 * (none_vectorstore.go:105-127).  Unknown ids are ignored like the Go `delete`. */
COLTT_API int coltt_b200_store_remove(coltt_store* s, const uint64_t* ids, size_t n);

/* Bulk ChangedVertex for rows that already live in device memory (shards too large to stage through host memory:
 * BASELINE config 4 is 10 M x 1536 fp32 = 61 GB per GPU): d_vecs is device fp32 [n][stride_floats] (stride_floats >=
 * dim), normalized + lowered exactly as coltt_b200_store_upsert does; row j of the call gets id = id_base + slot,
 * slot = rows already stored + j.  The store keeps no host id map for such rows: it answers searches, and rejects
 * upsert / remove / export / search_subset with COLTT_ERR_UNSUPPORTED. */
COLTT_API int coltt_b200_store_append_dev(coltt_store* s, const void* d_vecs, size_t n, uint32_t stride_floats, uint64_t id_base);

/* vectorspace.VertexSearch (edge/none_vectorstore.go:129-180 and the bf16/f16/f8 twins),
 * batched: nq un-normalized fp32 queries (host), top-k each.  The reference call is nq = 1;
 * `highCpu` has no meaning on the GPU.  Outputs (host, caller-allocated): out_ids and
 * out_scores are [nq][k], rows filled to out_counts[q] = min(k, rows), ascending score. */
COLTT_API int coltt_b200_store_search(coltt_store* s, const float* queries, size_t nq, int k, int select_mode,
                                      int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts);

/* vectorspace.FilterableVertexSearch (edge/none_vectorstore.go:182-253) given the candidate
 * ids that inverted.SearchWithExpression produced (pkg/inverted/search.go:113-119): the same
 * scan behind a gather front-end.  Ids not present are skipped like the Go map lookup. */
COLTT_API int coltt_b200_store_search_subset(coltt_store* s, const float* queries, size_t nq, const uint64_t* cand_ids,
                                             size_t n_cand, int k, int select_mode, uint64_t* out_ids,
                                             float* out_scores, int32_t* out_counts);

/* Device-resident variant for pipelines that keep queries and results on the GPU (multi-GPU
 * merge, benchmarks).  d_queries: device fp32 [nq][dim]; d_out: device coltt_hit [nq][k];
 * d_counts: device int32 [nq].  `stream` is a cudaStream_t (NULL = the store's own stream;
 * then the call returns after the work completed).  With a caller stream the call only
 * enqueues. */
typedef struct coltt_hit {
  uint64_t id;
  float score;
  uint32_t slot; /* row slot inside this store (diagnostic; ignore across stores) */
} coltt_hit;
COLTT_API int coltt_b200_store_search_dev(coltt_store* s, const void* d_queries, size_t nq, int k, int select_mode,
                                          int math_mode, void* d_out, void* d_counts, void* stream);

/* The exchange step of a sharded search (reference analogue: merging the 16 shard-local
 * queues, none_vectorstore.go:173-178): merges n_lists best-first lists of `k_in` hits per
 * query — d_lists is device coltt_hit [n_lists][nq][k_in] with d_list_counts int32
 * [n_lists][nq] — into device [nq][k] ascending, d_out_counts int32 [nq]. */
COLTT_API int coltt_b200_merge_topk_dev(int device, const void* d_lists, const void* d_list_counts, int n_lists,
                                        size_t nq, int k_in, int k, int select_mode, void* d_out, void* d_out_counts,
                                        void* stream);

/* Same merge for ONE packed message per shard (what the all-gather delivers): shard r's block starts at
 * d_packed + r*rank_stride_bytes and holds coltt_hit[nq][k_in], then at counts_offset_bytes int32 counts[nq]. */
COLTT_API int coltt_b200_merge_topk_dev2(int device, const void* d_packed, int n_lists, size_t nq, int k_in, int k,
                                         int select_mode, size_t rank_stride_bytes, size_t counts_offset_bytes, void* d_out,
                                         void* d_out_counts, void* stream);

/* vectorspace.SaveVertex / LoadVertex (edge/none_vectorstore.go:308-516; element widths
 * f16_vectorstore.go:338-343, f8_vectorstore.go:340): the reference's big-endian vertex blob.
 * Export writes metaCount = 0 for every vertex (metadata lives on the Go side); import skips
 * metadata records.  export: pass buf = NULL to query the size in *len. */
COLTT_API int coltt_b200_store_export(coltt_store* s, void* buf, size_t* len);
COLTT_API int coltt_b200_store_import(coltt_store* s, const void* buf, size_t len);

/* Stored (normalized + lowered) row for an id, as the reference keeps it in ENode.Vector:
 * dim elements of 4/2/1 bytes.  For tests and for Hnsw-style Get(). */
COLTT_API int coltt_b200_store_get_row(coltt_store* s, uint64_t id, void* out, size_t out_bytes);

/* ---- experimental CFLAT multi-vector search ------------------------------------------------
 * multiVectorVertex.MultiVertexSearch (experimental/multi_vector_vertex.go:85-137).  A CFLAT collection keeps one
 * fp32 vector per named field and vertex; here every field is a coltt_store (quant NONE, same dim / metric / device)
 * and the host applies each ChangedVertex / RemoveVertex to all of them in the same order, so they share the slot
 * layout.  The caller passes the INCLUDED query fields in request order: fields[j] the store of that field,
 * queries[j] its un-normalized fp32 query (dim floats, host), ratios[j] its Ratio (the Go side has already checked
 * that they sum to 100, experimental_analyzer.go:143-154).  score = sum_j scoreHelper(distance_j) * (float32(ratio_j)
 * / 100) in float32, unfused, in that order; the k largest scores come back in descending order (ties: id descending).
 * out_ids / out_scores hold k entries, *out_count = min(k, vertices). */
COLTT_API int coltt_b200_multi_search(coltt_store* const* fields, const float* const* queries, const int32_t* ratios,
                                      int n_fields, int k, uint64_t* out_ids, float* out_scores, int32_t* out_count);

/* ---- core/vectorindex HNSW ------------------------------------------------------------
 * Hnsw.Load (core/vectorindex/hnsw_commit.go:164-278): parses a Commit(header=true) blob
 * into a device-resident CSR graph + row matrix. */
COLTT_API int coltt_b200_hnsw_load(const void* commit_blob, size_t len, int device, coltt_hnsw** out);
COLTT_API void coltt_b200_hnsw_destroy(coltt_hnsw* h);
COLTT_API int coltt_b200_hnsw_len(coltt_hnsw* h, uint64_t* n);
/* Vector dimension of the index (hnswVertex.vector length): callers validate query length against it, as
 * Core.VectorSearch's callers do (core/core.go:633-695). */
COLTT_API int coltt_b200_hnsw_dim(coltt_hnsw* h, uint32_t* dim);
/* Hnsw.Search (hnsw.go:243-278), batched: normalize, greedy descent (hnsw.go:320-343),
 * searchLevel with ef = max(ef, k) (hnsw.go:345-389), trim to k, ascending.  ef <= 0 uses the
 * ef stored in the blob (hnsw_config.go:138).  Neighbour order = ascending id. */
COLTT_API int coltt_b200_hnsw_search(coltt_hnsw* h, const float* queries, size_t nq, int k, int ef, uint64_t* out_ids,
                                     float* out_scores, int32_t* out_counts);
/* Counters of the last search call: distance evaluations and expansions (for the roofline). */
COLTT_API int coltt_b200_hnsw_last_stats(coltt_hnsw* h, uint64_t* dist_evals, uint64_t* expansions);

/* Device time in milliseconds of the search kernel(s) of the last coltt_b200_hnsw_search call (CUDA events on its stream). */
COLTT_API int coltt_b200_hnsw_last_timing(coltt_hnsw* h, float* kernel_ms);

/* Bulk construction (replaces n x Hnsw.Insert, core/vectorindex/hnsw.go:104-167, for an initial load): per level,
 * the m nearest neighbours of every member by brute force on the tensor cores, exact fp32 edge distances, back
 * edges and pruneNeighbors (hnsw.go:449-474) — see csrc/hnsw_build.cu.  `levels` may be NULL (drawn as
 * RandomLevel does, hnsw.go:280-282, from `seed`) or hold the vertexLevel the Go side would pass to Insert.
 * Defaults as hnsw_config.go:135-162 when a field is <= 0: m 16 (mMax = m, mMax0 = 2m), ef 20, efConstruction 200.
 * The graph is a valid Hnsw (searchable here and, through coltt_b200_hnsw_commit, loadable by the Go side) but not
 * the one sequential insertion would produce. */
typedef struct coltt_hnsw_build_cfg {
  uint32_t dim;
  int32_t metric;          /* coltt_metric */
  int32_t m, ef, ef_construction;
  int32_t device;
  uint64_t seed;
} coltt_hnsw_build_cfg;
COLTT_API int coltt_b200_hnsw_build(const coltt_hnsw_build_cfg* cfg, const uint64_t* ids, const float* vecs /* n x dim, un-normalized */,
                                    const int32_t* levels /* nullable */, size_t n, coltt_hnsw** out);
/* Hnsw.Commit(w, header=true) (hnsw_commit.go:69-162): the reference's big-endian index blob; vertex metadata
 * count is written as 0 (metadata lives on the Go side).  Pass buf = NULL to query the size in *len. */
COLTT_API int coltt_b200_hnsw_commit(coltt_hnsw* h, void* buf, size_t* len);
/* Bulk-build wall times in ms: [0] ingest, [1] kNN search, [2] edge distances, [3] host graph assembly. */
COLTT_API int coltt_b200_hnsw_build_stats(coltt_hnsw* h, double* ms4, uint64_t* n_edges, int32_t* max_level);
/* [0] construction searches served by the tensor-core filter, [1] of those re-run exactly (margin not certified). */
COLTT_API int coltt_b200_hnsw_build_fast_stats(coltt_hnsw* h, uint64_t* out2);

/* ---- product quantization for the HNSW walk (BASELINE config 5) ------------------------------------------------------------
 * PARITY UNPINNED: the reference holds no PQ arithmetic (pkg/distancepq is dead code without any quantizer, pkg/hnswpq is
 * absent — SURVEY F5).  The parameters are the reference's ProductQuantizerParameters (pkg/models/hnsw_common.go:20-32):
 * NumCentroids in [2, 256], NumSubVectors >= 2 dividing the dimension, TriggerThreshold = rows the quantizer trains on
 * (sampled at a fixed stride; any value >= NumCentroids is accepted here).  Training is a Lloyd k-means per sub-vector on
 * the GPU; every vertex is then encoded as NumSubVectors code bytes. */
typedef struct coltt_pq_params {
  int32_t num_centroids;
  int32_t num_sub_vectors;
  int32_t trigger_threshold;
} coltt_pq_params;
COLTT_API int coltt_b200_hnsw_pq_train(coltt_hnsw* h, const coltt_pq_params* p, int iterations /* <= 0: 12 */);
/* Hnsw.Search with asymmetric distance computation over the codes (64 B per neighbour at dim 768 / 64 sub-vectors instead
 * of a 3 KB row).  rerank != 0: the ef survivors are re-scored with the reference's exact fp32 arithmetic, so returned
 * scores are true distances; rerank == 0: scores are the quantized estimates.  Judged on recall, not on bit parity. */
COLTT_API int coltt_b200_hnsw_pq_search(coltt_hnsw* h, const float* queries, size_t nq, int k, int ef, int rerank, uint64_t* out_ids,
                                        float* out_scores, int32_t* out_counts);

/* ---- sharded collections: one shard per GPU, ONE exchange of per-shard top-k lists, merge ----------------
 * Reference analogue: the 16 in-process shards of a vectorspace, each scanned into a shard-local queue and re-merged
 * (edge/none_vectorstore.go:152-178); rows -> shard by ShardVertex(id, 16) mod n_gpus (pkg/sharding/shard.go:34-41; the
 * host partitions the ids).  A coltt_comm is one rank: a GPU, its NCCL communicator, a stream and persistent exchange
 * buffers.  NCCL is loaded at run time (libnccl.so.2); without it these calls return COLTT_ERR_UNSUPPORTED. */
typedef struct coltt_comm coltt_comm;
/* One process driving n GPUs (the Go host): ncclCommInitAll over device_ids; out_comms receives n rank handles
 * (rank i = device_ids[i]).  n_dev = 1 needs no NCCL.  coltt_b200_shutdown destroys every handle created here. */
COLTT_API int coltt_b200_init(const int* device_ids, int n_dev, coltt_comm** out_comms);
COLTT_API void coltt_b200_shutdown(void);
/* One process per GPU: rank 0 makes the 128-byte rendezvous blob, the host hands it to every rank, each joins. */
COLTT_API int coltt_b200_comm_unique_id(void* out, size_t len);
COLTT_API int coltt_b200_comm_init_rank(const void* unique_id, int rank, int world, int device, coltt_comm** out);
/* With the peer-memory exchange the peers read this rank's exchange buffer while they merge: destroy a rank only after every
 * rank has returned from the last sharded search (a host-side barrier in a multi-process host; coltt_b200_shutdown and
 * coltt_b200_sharded_search_all take care of it inside one process). */
COLTT_API void coltt_b200_comm_destroy(coltt_comm* c);
COLTT_API int coltt_b200_comm_info(coltt_comm* c, int* rank, int* world, int* device);
/* How this rank exchanges the per-shard lists: COLTT_EXCHANGE_UNDECIDED before the first sharded search (the choice is a
 * collective vote taken there), COLTT_EXCHANGE_PEER = the all-gather is fused into the merge kernel, which reads every
 * peer's list out of the peer's HBM over NVLink (cudaIpc / peer access; every rank could map every peer),
 * COLTT_EXCHANGE_NCCL = ncclAllGather then merge (COLTT_P2P=0, > 8 ranks, or some pair without peer access). */
enum { COLTT_EXCHANGE_UNDECIDED = 0, COLTT_EXCHANGE_PEER = 1, COLTT_EXCHANGE_NCCL = 2 };
COLTT_API int coltt_b200_comm_exchange_mode(coltt_comm* c);
/* VertexSearch over the whole sharded collection, called by EVERY rank with the same queries / nq / k / modes (a
 * collective): local search of `shard` (hits written straight into the send buffer), ncclAllGather, merge.  Every rank
 * ends up with the merged answer; a rank may pass NULL outputs.  Host buffers; returns when the answer is in them. */
COLTT_API int coltt_b200_sharded_search(coltt_comm* c, coltt_store* shard, const float* queries, size_t nq, int k, int select_mode,
                                        int math_mode, uint64_t* out_ids, float* out_scores, int32_t* out_counts);
/* Device-resident form (d_queries fp32 [nq][dim], d_out coltt_hit [nq][k], d_counts int32 [nq]); with a caller stream the
 * call only enqueues (search, all-gather and merge are stream-ordered), with NULL it uses the rank's stream and waits. */
COLTT_API int coltt_b200_sharded_search_dev(coltt_comm* c, coltt_store* shard, const void* d_queries, size_t nq, int k, int select_mode,
                                            int math_mode, void* d_out, void* d_counts, void* stream);
/* All n ranks of this process in one call (one library thread per rank); outputs come from rank 0. */
COLTT_API int coltt_b200_sharded_search_all(coltt_comm* const* comms, coltt_store* const* shards, int n, const float* queries, size_t nq,
                                            int k, int select_mode, int math_mode, uint64_t* out_ids, float* out_scores,
                                            int32_t* out_counts);
/* Hnsw.Search over one independent sub-graph per GPU (each built over its row shard): local walk, the same all-gather,
 * merge nearest-first.  Collective like coltt_b200_sharded_search. */
COLTT_API int coltt_b200_sharded_hnsw_search(coltt_comm* c, coltt_hnsw* sub, const float* queries, size_t nq, int k, int ef,
                                             uint64_t* out_ids, float* out_scores, int32_t* out_counts);
/* The same with the product-quantized walk of every sub-graph (coltt_b200_hnsw_pq_search): BASELINE config 5. */
COLTT_API int coltt_b200_sharded_hnsw_pq_search(coltt_comm* c, coltt_hnsw* sub, const float* queries, size_t nq, int k, int ef, int rerank,
                                                uint64_t* out_ids, float* out_scores, int32_t* out_counts);

/* ---- timing (replaces pprof for this path) ---------------------------------
 * Device time in milliseconds of the kernels of the last search on this handle, measured
 * with CUDA events on the stream they ran on: [0] query prep, [1] scan/GEMM, [2] rerank,
 * [3] merge.  n = number of floats the caller provides. */
COLTT_API int coltt_b200_store_last_timing(coltt_store* s, float* ms, int n);
/* Per-phase events are recorded only while timing is on (default off: the event records sit between
 * kernels of a sub-millisecond search). */
COLTT_API int coltt_b200_store_set_timing(coltt_store* s, int on);
/* COLTT_MATH_FAST statistics of a handle: out2[0] = queries answered through the tensor-core filter, out2[1] = how many of
 * them the certificate sent to the exact re-run. */
COLTT_API int coltt_b200_store_fast_stats(coltt_store* s, uint64_t* out2);
/* The certificate margin COLTT_MATH_FAST uses for a collection of this dimension, relative to ||q|| * ||row||
 * (DESIGN.md section 5): tests compare the filter's raw scores against it. */
COLTT_API float coltt_b200_fast_eps_rel(uint32_t dim);
/* Kernels this library has launched in this process so far (bench.py's gpu_launches). */
COLTT_API uint64_t coltt_b200_kernel_launches(void);

/* ---- page-locked host buffers ------------------------------------------------------------------------------------------
 * Every host-pointer search entry point (coltt_b200_store_search, coltt_b200_sharded_search, ...) checks whether the
 * caller's query buffer is page-locked memory known to CUDA: if so it is the DMA source itself; otherwise the queries are
 * first copied into the handle's own pinned staging buffer (one host memcpy of nq * dim * 4 bytes: ~70 us of a 0.5 ms
 * batch at config 2).  A host that assembles its batches in a buffer from coltt_b200_host_alloc (the micro-batcher,
 * INTEGRATION.md) therefore skips that copy.  The library never keeps the pointer past the call.  Needs a GPU. */
COLTT_API int coltt_b200_host_alloc(size_t bytes, void** out);
COLTT_API void coltt_b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* COLTT_B200_H */
