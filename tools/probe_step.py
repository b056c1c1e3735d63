"""Host-vs-device cost of one search step at config 2 (1 M x 768 fp16, batch 256, top-10, FAST):
enqueue wall time per step with the stream left to run ahead (host cost) against the CUDA-event step time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coltt_b200 as cb
from coltt_b200 import _lib

n, d, nq, k = int(os.environ.get("N", 1_000_000)), 768, int(os.environ.get("NQ", 256)), 10
g = np.random.Generator(np.random.Philox(1))
rows = np.empty((n, d), np.float32)
for i in range(0, n, 100_000):
    rows[i:i + 100_000] = g.standard_normal((min(100_000, n - i), d), dtype=np.float32)
sp = cb.VectorSpace("p", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_BF16), capacity_hint=n, select_mode=cb.SELECT_NEAREST, math_mode=cb.MATH_FAST)
sp.ChangedVertices(np.arange(n, dtype=np.uint64) + 1, rows)
dev = torch.device("cuda", 0)
q = torch.from_numpy(g.standard_normal((nq, d), dtype=np.float32)).to(dev)
out = torch.zeros((nq, k, 4), dtype=torch.int32, device=dev)
cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
L = _lib.lib()
qp, op, cp, sh = q.data_ptr(), out.data_ptr(), cnt.data_ptr(), st.cuda_stream


def step():
    _lib.check(L.coltt_b200_store_search_dev(sp._h, qp, nq, k, cb.SELECT_NEAREST, cb.MATH_FAST, op, cp, sh))


for _ in range(20):
    step()
torch.cuda.synchronize()
for steps in (50, 400):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    for _ in range(steps):
        step()
    e1.record(st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"steps={steps}: host enqueue {1e3 * (t1 - t0) / steps:.4f} ms/step, device {e0.elapsed_time(e1) / steps:.4f} ms/step, "
          f"wall to drain {1e3 * (t2 - t0) / steps:.4f} ms/step", flush=True)
