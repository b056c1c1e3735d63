"""HNSW search probe (BASELINE config 3 shape at an oracle-buildable N): device time, distance evaluations,
achieved gather GB/s, next to the oracle (reference restatement, 1 thread) on the same host."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coltt_b200 as cb
from oracle import oracle as orc
n, d, k, ef, nq = int(os.environ.get("N", 8000)), 768, 10, 128, 256
g = np.random.Generator(np.random.Philox(3))
vecs = g.standard_normal((n, d), dtype=np.float32)
ids = np.arange(1, n + 1, dtype=np.uint64)
orc.use_reference_kernels(True)
t0 = time.perf_counter(); h = orc.Hnsw(d, orc.COSINE); h.build(ids, vecs); tb = time.perf_counter() - t0
gpu = cb.Hnsw.Load(h.commit())
qs = g.standard_normal((nq, d), dtype=np.float32)
h.set_ef(ef); h.stats(reset=True)
t0 = time.perf_counter()
for q in qs[:32]:
    h.search(q, k)
tc = (time.perf_counter() - t0) / 32
for _ in range(3):
    gpu.BatchSearch(qs, k, ef)
t0 = time.perf_counter(); reps = 10
for _ in range(reps):
    gpu.BatchSearch(qs, k, ef)
tg = (time.perf_counter() - t0) / reps
st = gpu.last_stats()
byts = st["dist_evals"] * (d * 4 + 8) + st["expansions"] * 32 * 4
print(f"hnsw N={n} d={d} ef={ef} k={k}: oracle build {tb:.1f}s; CPU {tc*1e3:.3f} ms/query (1 thread); GPU batch of {nq}: {tg*1e3:.3f} ms wall "
      f"= {nq/tg:.0f} QPS; evals/query {st['dist_evals']/nq:.0f}, expansions/query {st['expansions']/nq:.0f}; gather {byts/tg/1e9:.1f} GB/s")
