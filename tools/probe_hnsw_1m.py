"""Config 3 probe (core/vectorindex HNSW fp32 dim=768 N=1M efSearch=128 top-10, one B200): bulk-build the graph on
the GPU (csrc/hnsw_build.cu), search batches of 256, report build time, QPS, distance evaluations per query,
achieved random-gather bandwidth (E*(dim*4+8) + X*mMax0*4 bytes) and recall@10 against exact search."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coltt_b200 as cb

n, d = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 768))
ef, k = int(os.environ.get("EF", 128)), 10
nqs = [int(x) for x in os.environ.get("NQ", "256").split(",")]
data = os.environ.get("DATA", "iid")   # iid: N(0,1)^d (no structure: the worst case for any graph index);
g = np.random.Generator(np.random.Philox(0xC0177))       # latent: 32-d latent through a random projection + 10% noise
rows = np.empty((n, d), np.float32)
LAT = 32
A = np.random.Generator(np.random.Philox(0xA)).standard_normal((LAT, d), dtype=np.float32)


def gen(gen_, m):
    if data == "iid":
        return gen_.standard_normal((m, d), dtype=np.float32)
    return (gen_.standard_normal((m, LAT), dtype=np.float32) @ A + np.float32(0.1) * gen_.standard_normal((m, d), dtype=np.float32)).astype(np.float32)


for i in range(0, n, 100_000):
    rows[i:i + 100_000] = gen(g, min(100_000, n - i))
ids = np.arange(n, dtype=np.uint64) + 1
t0 = time.perf_counter()
h = cb.Hnsw.Build(ids, rows, metric=cb.Distance_Cosine, m=16, ef=ef)
t_build = time.perf_counter() - t0
bs = h.build_stats()
qs_all = gen(np.random.Generator(np.random.Philox(0xC0178)), max(nqs))
# exact ground truth: FLAT fp32 store, NEAREST (bit-identical to the reference arithmetic)
sp = cb.VectorSpace("gt", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None), capacity_hint=n, select_mode=cb.SELECT_NEAREST)
sp.ChangedVertices(ids, rows)
rq = min(64, min(nqs))
wi, ws, wc = sp.BatchVertexSearch(qs_all[:rq], k, math_mode=cb.MATH_EXACT)
sp.close()
for nq in nqs:
    qs = qs_all[:nq]
    for _ in range(2):
        h.BatchSearch(qs, k, ef)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        gi, gs, gc = h.BatchSearch(qs, k, ef)
    wall = (time.perf_counter() - t0) / reps
    st = h.last_stats()
    E, X = st["dist_evals"] / nq, st["expansions"] / nq
    gather_bytes = st["dist_evals"] * (d * 4 + 8) + st["expansions"] * 32 * 4
    rec = np.mean([len(set(wi[j, :k].tolist()) & set(gi[j, :k].tolist())) / k for j in range(rq)])
    print(json.dumps({"probe": "hnsw_c3", "data": data, "n": n, "dim": d, "ef": ef, "k": k, "batch": nq, "build_s": round(t_build, 2), "build_stats": bs,
                      "search_ms_per_batch": round(wall * 1e3, 3), "qps": round(nq / wall), "dist_evals_per_query": round(E, 1),
                      "expansions_per_query": round(X, 1), "gather_GBps": round(gather_bytes / wall / 1e9, 1), "recall_at_10": float(rec),
                      "recall_queries": rq}), flush=True)
