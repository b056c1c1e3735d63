// multi.cu — CFLAT multi-vector search.
//
// Replaces experimental.multiVectorVertex.MultiVertexSearch (experimental/multi_vector_vertex.go:85-137): every
// vertex holds one fp32 vector per named field (all of the collection's dimension, normalized at ChangedVertex
// for cosine, :60-67; CFLAT supports no quantization, mutli_vecspace.go:60-64); a query brings one vector and a
// ratio per included field and the score of a vertex is
//     score = 0; for each included field, in request order:
//         score += scoreHelper(Distance(vertex[field], Normalize(query[field]))) * (float32(ratio) / 100)
// in float32 without fusion; the topK LARGEST scores are returned in descending order (multi_priority_queue.go:47-75:
// min-queue, pop when over capacity, then sort descending).
//
// Here a field is an ordinary fp32 store (store.cu) and all field stores of a collection share the slot layout
// (the host applies the same upsert / remove sequence to each).  One K1 launch per included field streams that
// field's rows once (HBM-bound, `flat_scan.cu`) and adds its term to a per-slot running score (4 B per row and
// field next to dim*4 B of row data); the last field's launch feeds the sum straight into the fused warp top-K
// (select = largest), and K5 merges.  Algorithmic bytes: F * n * (dim*4 + 4 + 8).
// Ties: the reference's order among equal scores is Go-heap / map-iteration dependent; here it is the reverse of
// the total order T (score descending, then id descending) — the rule the CPU checker applies as well.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <vector>

#include "kernels.cuh"
#include "store.h"

namespace coltt {

static int multi_search(Store* const* fields, const float* const* queries, const int32_t* ratios, int n_fields, int k, uint64_t* out_ids,
                        float* out_scores, int32_t* out_count) {
  if (!fields || !queries || !ratios || !out_ids || !out_scores || !out_count) return fail(COLTT_ERR_INVALID, "null argument");
  if (n_fields <= 0) return fail(COLTT_ERR_INVALID, "no included vector field");
  if (k <= 0) return fail(COLTT_ERR_INVALID, "top-k must be positive");
  Store* s0 = fields[0];
  for (int j = 0; j < n_fields; j++) {
    Store* s = fields[j];
    if (!s || !queries[j]) return fail(COLTT_ERR_INVALID, "null field store or query");
    if (s->elem != ELEM_F32) return fail(COLTT_ERR_UNSUPPORTED, "not support quantization type");   // mutli_vecspace.go:63
    if (s->dim != s0->dim || s->cfg.metric != s0->cfg.metric || s->device != s0->device)
      return fail(COLTT_ERR_INVALID, "field stores of one collection must share dim, distance and device");
  }
  // shared locks on every distinct field store, in address order
  std::vector<Store*> uniq(fields, fields + n_fields);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::vector<std::shared_lock<std::shared_mutex>> locks;
  for (Store* s : uniq) locks.emplace_back(s->mu);
  const size_t n = s0->n_rows;
  for (Store* s : uniq)
    if (s->n_rows != n) return fail(COLTT_ERR_INVALID, "field stores of one collection must hold the same vertices");
  *out_count = 0;
  if (n == 0) return COLTT_OK;
  COLTT_CUDA(cudaSetDevice(s0->device));
  auto ctx = s0->acquire_ctx(nullptr);
  if (!ctx) return fail(COLTT_ERR_CUDA, "could not create a search context");
  struct Rel { Store* s; std::unique_ptr<SearchCtx>* c; ~Rel() { s->release_ctx(std::move(*c)); } } rel{s0, &ctx};
  SearchCtx& c = *ctx;
  cudaStream_t st = c.stream;
  const uint32_t dim = s0->dim, q_stride = (dim + 7) / 8 * 8;
  const uint32_t k_eff = (uint32_t)std::min<size_t>((size_t)k, n);
  ScanPlan plan;
  int rc = plan_flat_scan(ELEM_F32, dim, s0->row_stride, (uint32_t)n, 1, k_eff, s0->n_sms, &plan);
  if (rc) return rc;
  if ((rc = c.q_in.ensure((size_t)n_fields * dim * 4)) || (rc = c.q_deq.ensure((size_t)n_fields * q_stride * 4)) ||
      (rc = c.q_n2.ensure((size_t)n_fields * 4)) || (rc = c.warp_lists.ensure(plan.warp_list_bytes)) ||
      (rc = c.cta_lists.ensure(plan.cta_list_bytes)) || (rc = c.cta_counts.ensure(plan.cta_count_bytes)) ||
      (rc = c.multi_acc.ensure(s0->capacity * 4)) || (rc = c.out.ensure((size_t)k_eff * sizeof(Hit))) || (rc = c.counts.ensure(4)) ||
      (rc = c.h_q.ensure((size_t)n_fields * dim * 4)) || (rc = c.h_out.ensure((size_t)k_eff * sizeof(Hit) + 4)))
    return rc;
  for (int j = 0; j < n_fields; j++) std::memcpy((float*)c.h_q.p + (size_t)j * dim, queries[j], (size_t)dim * 4);
  COLTT_CUDA(cudaMemcpyAsync(c.q_in.p, c.h_q.p, (size_t)n_fields * dim * 4, cudaMemcpyHostToDevice, st));
  // Normalize(vectors.GetVector()) for cosine (multi_vector_vertex.go:97-101): all field queries in one launch
  PrepParams pp{};
  pp.in = (const float*)c.q_in.p; pp.n = (size_t)n_fields; pp.in_stride = dim; pp.dim = dim; pp.smem_stride = (dim + 3) / 4 * 4;
  pp.normalize = s0->cfg.metric == COLTT_COSINE;
  pp.norm2_out = (float*)c.q_n2.p; pp.norm2_by_slot = 0;
  pp.deq_out = (float*)c.q_deq.p; pp.deq_stride = q_stride;
  if ((rc = launch_prep_rows(pp, ELEM_F32, st))) return rc;
  for (int j = 0; j < n_fields; j++) {
    Store* s = fields[j];
    ScanParams sp{};
    sp.rows = s->d_rows; sp.row_stride = s->row_stride; sp.dim = dim; sp.n_items = (uint32_t)n; sp.subset = nullptr;
    sp.row_norm2 = s->d_norm2; sp.ids = s->d_ids;
    sp.queries = (const float*)c.q_deq.p + (size_t)j * q_stride; sp.q_norm2 = (const float*)c.q_n2.p + j; sp.q_stride = q_stride;
    sp.nq = 1; sp.k = k_eff; sp.nearest = 0; sp.metric = s0->cfg.metric;
    sp.warp_lists = (Hit*)c.warp_lists.p; sp.cta_lists = (Hit*)c.cta_lists.p; sp.cta_counts = (int*)c.cta_counts.p;
    sp.multi_acc = (float*)c.multi_acc.p;
    sp.multi_w = (float)(uint32_t)ratios[j] / 100.0f;     // float32(vectors.Ratio) / 100
    sp.multi_first = j == 0; sp.multi_last = j == n_fields - 1;
    if ((rc = launch_flat_scan(sp, plan, ELEM_F32, st))) return rc;
  }
  MergeParams mp{};
  mp.lists = (const Hit*)c.cta_lists.p; mp.counts = (const int*)c.cta_counts.p; mp.n_lists = plan.grid_x;
  mp.nq = 1; mp.k_in = k_eff; mp.k = k_eff; mp.nearest = 0; mp.in_best_first = 1;
  mp.out = (Hit*)c.out.p; mp.out_counts = (int*)c.counts.p;
  if ((rc = launch_merge_topk(mp, st))) return rc;
  Hit* h_hits = (Hit*)c.h_out.p;
  int* h_cnt = (int*)((uint8_t*)c.h_out.p + (size_t)k_eff * sizeof(Hit));
  COLTT_CUDA(cudaMemcpyAsync(h_hits, c.out.p, (size_t)k_eff * sizeof(Hit), cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaMemcpyAsync(h_cnt, c.counts.p, 4, cudaMemcpyDeviceToHost, st));
  COLTT_CUDA(cudaStreamSynchronize(st));
  c.used = true;
  c.last_stream = st;
  cudaEventRecord(c.done, st);
  // K5 reports the K largest in T order (ascending): sort.Slice(... Score > ...) wants them descending
  const int cnt = *h_cnt;
  *out_count = cnt;
  for (int i = 0; i < cnt; i++) {
    out_ids[i] = h_hits[cnt - 1 - i].id;
    out_scores[i] = h_hits[cnt - 1 - i].score;
  }
  return COLTT_OK;
}

}  // namespace coltt

extern "C" {
COLTT_API int coltt_b200_multi_search(coltt_store* const* fields, const float* const* queries, const int32_t* ratios, int n_fields, int k,
                                      uint64_t* out_ids, float* out_scores, int32_t* out_count) {
  return coltt::multi_search(reinterpret_cast<coltt::Store* const*>(fields), queries, ratios, n_fields, k, out_ids, out_scores, out_count);
}
}
