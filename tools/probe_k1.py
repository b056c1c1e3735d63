"""Single-query exact scan (the reference's actual call shape) for ncu: 1M x 768, fp16 or fp32 rows."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coltt_b200 as cb
quant = {"fp16": cb.Quantization_BF16, "fp32": cb.Quantization_None}[os.environ.get("Q", "fp16")]
n, d = int(os.environ.get("N", 1_000_000)), 768
g = np.random.Generator(np.random.Philox(1))
rows = np.empty((n, d), np.float32)
for i in range(0, n, 100_000):
    rows[i:i + 100_000] = g.standard_normal((min(100_000, n - i), d), dtype=np.float32)
sp = cb.VectorSpace("p", cb.Metadata(d, cb.Distance_Cosine, quant), capacity_hint=n, select_mode=cb.SELECT_NEAREST)
sp.set_timing(True)
sp.ChangedVertices(np.arange(n, dtype=np.uint64) + 1, rows)
q = g.standard_normal((1, d), dtype=np.float32)
for _ in range(4):
    sp.BatchVertexSearch(q, 10, math_mode=cb.MATH_EXACT)
print(sp.last_timing_ms())
