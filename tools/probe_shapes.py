"""Device-time probe of the search kernels at the reference's call shapes (not a bench line):
prints last_timing_ms() for (store type, batch, math) at N rows x dim."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coltt_b200 as cb

n, d = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 768))
g = np.random.Generator(np.random.Philox(1))
rows = np.empty((n, d), np.float32)
for i in range(0, n, 100_000):
    rows[i:i + 100_000] = g.standard_normal((min(100_000, n - i), d), dtype=np.float32)
ids = np.arange(n, dtype=np.uint64) + 1
for quant, name in ((cb.Quantization_BF16, "fp16"), (cb.Quantization_None, "fp32"), (cb.Quantization_F8, "f8c")):
    sp = cb.VectorSpace("p", cb.Metadata(d, cb.Distance_Cosine, quant), capacity_hint=n, select_mode=cb.SELECT_NEAREST)
    sp.set_timing(True)
    sp.ChangedVertices(ids, rows)
    es = {cb.Quantization_BF16: 2, cb.Quantization_None: 4, cb.Quantization_F8: 1}[quant]
    for nq in (1, 8, 64, 128, 256):
        for math, mname in ((cb.MATH_EXACT, "exact"), (cb.MATH_FAST, "fast")):
            if math == cb.MATH_FAST and quant != cb.Quantization_BF16:
                continue
            if math == cb.MATH_EXACT and nq > 8 and quant != cb.Quantization_BF16:
                continue
            qs = g.standard_normal((nq, d), dtype=np.float32)
            for _ in range(3):
                sp.BatchVertexSearch(qs, 10, math_mode=math)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                sp.BatchVertexSearch(qs, 10, math_mode=math)
            wall = (time.perf_counter() - t0) / reps * 1e3
            t = sp.last_timing_ms()
            gbs = n * d * es / (t["scan"] * 1e-3) / 1e9 if t["scan"] > 0 else 0
            print(f"{name} nq={nq:4d} {mname:5s} scan={t['scan']:.3f}ms prep={t['prep']:.3f} post={t['merge']:.3f} wall={wall:.3f}ms  scan_GB/s(one pass)={gbs:.0f}", flush=True)
    sp.close()
