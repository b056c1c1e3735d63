"""K2 (tcgen05 filter) probe: device time of the FAST path at N rows (env N, default 1 M) for a batch of 256,
with whatever COLTT_DEBUG_* / COLTT_FAST_* knobs are set in the environment (role timers print to stderr)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coltt_b200 as cb
n, d, nq = int(os.environ.get("N", 1_000_000)), 768, int(os.environ.get("NQ", 256))
g = np.random.Generator(np.random.Philox(1))
rows = np.empty((n, d), np.float32)
for i in range(0, n, 100_000):
    rows[i:i + 100_000] = g.standard_normal((min(100_000, n - i), d), dtype=np.float32)
sp = cb.VectorSpace("p", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_BF16), capacity_hint=n, select_mode=cb.SELECT_NEAREST)
sp.set_timing(True)
sp.ChangedVertices(np.arange(n, dtype=np.uint64) + 1, rows)
qs = g.standard_normal((nq, d), dtype=np.float32)
ts = []
for i in range(12):
    sp.BatchVertexSearch(qs, 10, math_mode=cb.MATH_FAST)
    ts.append(sp.last_timing_ms()["scan"])
ts = sorted(ts[4:])
print(f"K2 N={n} nq={nq} flags={os.environ.get('COLTT_DEBUG_FLAGS','0')} ns={os.environ.get('COLTT_FAST_NS','max')}: scan median {ts[len(ts)//2]:.4f} ms min {ts[0]:.4f}  "
      f"({n*d*2/ts[len(ts)//2]/1e6:.0f} GB/s)", flush=True)
