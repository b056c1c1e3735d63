"""Micro-batcher: concurrent single-query callers are coalesced into batched searches and each gets
exactly its own row back.  Host logic only — the batched search here is a numpy stand-in."""
import threading

import numpy as np
import pytest

from coltt_b200.batcher import MicroBatcher


def _brute(rows):
    calls = []

    def search(qs, k):
        calls.append(len(qs))
        d = ((qs[:, None, :] - rows[None, :, :]) ** 2).sum(-1)
        idx = np.argsort(d, axis=1, kind="stable")[:, :k]
        return idx.astype(np.uint64), np.take_along_axis(d, idx, 1).astype(np.float32), np.full(len(qs), min(k, len(rows)), np.int32)
    return search, calls


def test_concurrent_callers_are_coalesced_and_get_their_own_rows():
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((500, 16)).astype(np.float32)
    search, calls = _brute(rows)
    mb = MicroBatcher(search, 16, max_batch=32, max_wait_ms=20.0)
    qs = rng.standard_normal((200, 16)).astype(np.float32)
    out = [None] * len(qs)

    def worker(lo, hi):
        for i in range(lo, hi):
            out[i] = mb.VertexSearch(qs[i], 5)
    ths = [threading.Thread(target=worker, args=(i * 25, (i + 1) * 25)) for i in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    mb.close()
    want_ids, want_sc, _ = search(qs, 5)
    for i in range(len(qs)):
        assert np.array_equal(out[i][0], want_ids[i]) and out[i][1].tobytes() == want_sc[i].tobytes()
    assert mb.served == 200 and mb.batches < 200 and max(calls[:-1]) <= 32 and max(calls[:-1]) > 1


def test_mixed_topk_flush_on_timeout_errors_and_close():
    rows = np.eye(8, dtype=np.float32)
    search, calls = _brute(rows)
    mb = MicroBatcher(search, 8, max_batch=256, max_wait_ms=250.0)   # long enough for the three submits below to be queued together
    f1, f2, f3 = mb.submit(rows[1], 3), mb.submit(rows[2], 1), mb.submit(rows[3], 3)
    assert f1.result(5)[0][0] == 1 and f2.result(5)[0].tolist() == [2] and f3.result(5)[0][0] == 3
    assert calls[:2] == [2, 1]                       # equal-topK requests ride together, the odd one alone
    with pytest.raises(ValueError, match="Dim Length"):
        mb.submit(np.zeros(5, np.float32), 3)
    mb.close()
    with pytest.raises(RuntimeError):
        mb.submit(rows[0], 1)

    def boom(qs, k):
        raise RuntimeError("device lost")
    mb2 = MicroBatcher(boom, 8, max_wait_ms=0.1)
    with pytest.raises(RuntimeError, match="device lost"):
        mb2.VertexSearch(rows[0], 2)
    mb2.close()


def test_bench_roofline_picks_the_binding_floor():
    """bench.py's roofline: both floors from the measured peaks, the larger one binds (config 2 sits above the ridge)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    peaks = {"hbm_gbs": 6541.1, "tf_burst": 1624.4, "tf_sustained": 1364.4, "src": "measured"}
    n, d, nq, k = 1_000_000, 768, 256, 10
    alg = n * d * 2 + n * 4 + nq * d * 4 + nq * k * 16
    # a 0.36 ms kernel inside a 0.42 ms step is judged against the BURST cuBLAS figure
    r = bench.roofline_fast(alg, 2.0 * nq * n * d, 0.3614, 0.42, peaks, 1597533448, "gemm_filter_pair_kernel", False)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == 1624.4 and abs(r["achieved"] / r["peak"] - 0.670) < 0.01
    assert abs(r["tensor_frac_vs_sustained"] - 0.797) < 0.01
    assert abs(r["hbm_frac"] - 0.652) < 0.01 and r["floor_ms"]["tensor"] > r["floor_ms"]["hbm"]
    r1 = bench.roofline_fast(alg, 2.0 * 8 * n * d, 0.3, 0.35, peaks, None, "gemm_filter_kernel", False)     # 8 queries: HBM-bound
    assert r1["bound"] == "hbm" and r1["unit"] == "GB/s" and r1["peak"] == 6541.1
    # config 4: a 13 ms fp8 kernel is judged against 2 x the SUSTAINED bf16 figure
    n4, d4, q4 = 10_000_000, 1536, 1024
    r4 = bench.roofline_fast(n4 * d4, 2.0 * q4 * n4 * d4, 13.0, 14.0, peaks, None, "gemm_filter_pair_kernel", True)
    assert r4["bound"] == "tensor" and r4["peak"] == 2 * 1364.4
    for key in ("bound", "achieved", "peak", "unit", "traffic"):
        assert key in r and key in r1


def test_bench_clock_sampler_summarises_only_rows_inside_the_load_window():
    """bench.py's nvidia-smi sampler time-stamps its rows; stop(t_from, t_to) must ignore rows printed before the load began
    (idle clocks) and keep throttle reasons of the rows inside the window."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    row = "0, {sm}, 1965, 700.0, 0x0, Not Active, Not Active, Not Active, {cap}"
    s.rows = [(10.00, row.format(sm=345, cap="Not Active")),       # idle, before the load
              (10.30, row.format(sm=1965, cap="Not Active")),
              (10.35, row.format(sm=1500, cap="Active")),
              (10.40, row.format(sm=1440, cap="Active")),
              (11.00, row.format(sm=345, cap="Not Active")),       # long after
              (10.36, "garbage line")]
    c = s.stop(10.25, 10.40)
    assert c["samples"] == 3 and c["sm_mhz"] == 1500.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    assert s2.stop()["reasons"] == ["nvidia-smi unavailable"]
