#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: QPS @ recall@10 of edge FLAT search.

Workload (BASELINE.json configs[1]): edge FLAT "bf16" (= IEEE fp16 rows) cosine,
dim 768, 1M vectors per GPU shard, batch 256, top-10, synthetic N(0,1) data (seeds 0xC0177 / 0xC0178).

One "step" = one batch of 256 queries answered over the rank's 1M-row shard (and, at N>1, merged
across shards after one all-gather of the per-shard top-k).  Timed with CUDA events on the launching
stream, barrier + synchronize on both sides, max over ranks.  `value` has the queries resident in
HBM; `e2e` goes through the host-pointer C-ABI call (H2D of the queries and D2H of the results inside
the timed region).  The shard (1.54 GB) is >12x the 126 MB L2, so every step streams it from HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
    python bench.py --impl reference ...                            the reference's CPU path (oracle arm)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASE_SEED, QUERY_SEED = 0xC0177, 0xC0178
METRIC_NAME = "QPS @ recall@10, dim=768, 1M vecs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flat", choices=["flat", "hnsw", "c4", "c5"],
                    help="flat = BASELINE configs[1] (the headline, default); hnsw = configs[2] (core/vectorindex HNSW, one GPU); "
                         "c4 = configs[3] (edge FLAT fp8 cosine dim=1536 N=10M/shard batch=1024 top-100, E4M3 store); "
                         "c5 = configs[4] (HNSW + PQ dim=768 N=10M efSearch=256 sharded 4 GPUs: 2.5M vertices per GPU)")
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--quant", default=None, choices=["bf16", "none", "f8_e4m3"],
                    help="store quantization (default: the workload's); none = fp32 rows, filtered through their fp16 shadow")
    ap.add_argument("--no-extras", action="store_true", help="skip the compact c3 / c4 records of the default line")
    ap.add_argument("--math", default="auto", choices=["auto", "exact", "fast"])
    ap.add_argument("--cpu-rows", type=int, default=250_000, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-queries", type=int, default=8)
    ap.add_argument("--recall-queries", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / recall legs (profiling runs)")
    a = ap.parse_args()
    wl = WORKLOADS["flat" if a.workload == "hnsw" else a.workload]
    if a.workload == "c5" and a.ef == 128:
        a.ef = 256
    for key in ("rows", "dim", "batch", "k", "steps", "warmup", "quant"):
        if getattr(a, key) is None:
            setattr(a, key, wl[key])
    return a


# The two FLAT workloads.  `device_gen`: rows are generated on the GPU (10 M x 1536 fp32 = 61 GB does not fit host memory)
WORKLOADS = {
    "flat": dict(rows=1_000_000, dim=768, batch=256, k=10, steps=1000, warmup=10, quant="bf16", es=2, device_gen=False,
                 name="edge FLAT bf16(=fp16) cosine dim=768 N=1M/GPU batch=256 top-10 (BASELINE configs[1])",
                 dtype="f16 rows/queries, f32 accumulate", metric="QPS @ recall@10, dim=768, 1M vecs"),
    "c4": dict(rows=10_000_000, dim=1536, batch=1024, k=100, steps=20, warmup=3, quant="f8_e4m3", es=1, device_gen=True,
               name="edge FLAT f8 (E4M3 store) cosine dim=1536 N=10M/GPU batch=1024 top-100 (BASELINE configs[3])",
               dtype="e4m3 rows/queries, f32 accumulate", metric="QPS @ recall@100, dim=1536, 10M vecs/shard"),
    "c5": dict(rows=2_500_000, dim=768, batch=1024, k=10, steps=10, warmup=3, quant="pq", es=1, device_gen=False,
               name="HNSW + PQ (64 sub-vectors x 256 centroids) cosine dim=768 N=10M over 4 GPUs (2.5M/GPU) efSearch=256 top-10 (BASELINE configs[4])",
               dtype="u8 PQ codes (ADC, f32 tables) + f32 exact re-rank", metric="QPS @ recall@10, dim=768, HNSW+PQ, 2.5M vecs/shard"),
}


def gen_rows(n, d, seed, chunk=100_000):
    out = np.empty((n, d), dtype=np.float32)
    g = np.random.Generator(np.random.Philox(seed))
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        out[i:i + m] = g.standard_normal((m, d), dtype=np.float32)
    return out


def shard_ids(n, rank):
    # dense-ish unique u64 ids, disjoint across ranks
    return (np.arange(n, dtype=np.uint64) * np.uint64(8) + np.uint64(rank)) + np.uint64(1 << 32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_from=None, t_to=None):
        """Summary of the samples taken in [t_from, t_to] (perf_counter times; default: all of them)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_row, r in self.rows:
            # a row printed at t_row describes the ~50 ms before it
            if (t_from is not None and t_row < t_from) or (t_to is not None and t_row > t_to + 0.06):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
        except Exception:
            v = None
        if v is not None:
            return v
    return None


def measured_peaks():
    """Denominators of the roofline: the driver-written MEASURED_PEAKS.json (HBM copy GB/s; cuBLAS dense bf16 TFLOP/s, burst =
    best single call, sustained = back to back for seconds), else the fallback B200_PROFILING.md states."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        burst = j.get("bf16_tflops", 1590.0)
        return {"hbm_gbs": j.get("hbm_gbs", 6650.0), "tf_burst": burst, "tf_sustained": j.get("bf16_tflops_sustained", burst), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1590.0, "src": "fallback"}


def roofline_fast(alg_bytes, flops, kern_ms, step_ms, peaks, traffic, kname, fp8):
    """Roofline of the tcgen05 filter kernel.  Both floors are computed from the MEASURED peaks; the larger one binds.
    Tensor peak: a kernel that is most of a short step (config 2: 0.3 ms of a 0.4 ms step, clocks at max, no power cap) is
    judged against the BURST cuBLAS figure; a kernel that runs for milliseconds back to back (config 4) against the
    SUSTAINED one.  fp8 (kind::f8f6f4) runs at twice the bf16 rate on this tensor core; MEASURED_PEAKS has only bf16, so
    the fp8 denominator is 2 x the measured bf16 figure (stated in `tensor_peak_note`)."""
    sec = kern_ms / 1e3
    long_kernel = kern_ms >= 2.0
    tf = (peaks["tf_sustained"] if long_kernel else peaks["tf_burst"]) * (2.0 if fp8 else 1.0)
    gbs, tfs = alg_bytes / sec / 1e9, flops / sec / 1e12
    t_hbm, t_tensor = alg_bytes / (peaks["hbm_gbs"] * 1e9), flops / (tf * 1e12)
    roof = {"kernel": kname, "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops, "traffic": traffic,
            "hbm_gbs": gbs, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_frac": gbs / peaks["hbm_gbs"],
            "tensor_tflops": tfs, "tensor_peak_tflops": tf, "tensor_frac": tfs / tf,
            "tensor_peak_note": ("2 x " if fp8 else "") + ("bf16_tflops_sustained" if long_kernel else "bf16_tflops (burst)") + " of MEASURED_PEAKS.json",
            "tensor_frac_vs_burst": tfs / (peaks["tf_burst"] * (2.0 if fp8 else 1.0)),
            "tensor_frac_vs_sustained": tfs / (peaks["tf_sustained"] * (2.0 if fp8 else 1.0)),
            "floor_ms": {"hbm": t_hbm * 1e3, "tensor": t_tensor * 1e3}, "kernel_share_of_step": kern_ms / step_ms if step_ms else None}
    if t_tensor >= t_hbm:
        roof.update({"bound": "tensor", "achieved": tfs, "peak": tf, "unit": "TFLOP/s"})
    else:
        roof.update({"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s"})
    return roof


# ------------------------------------------------------------------------------- reference arm
def cpu_store(wl, dim, rows, ids):
    """The reference's CPU store for this workload, through the oracle: "bf16" -> bf16_vectorstore.go (fp16 codec); the f8
    workload -> f8_vectorstore.go with the reference's own (literal) f8 codec — the CPU cost per row (dequantise both operands,
    one AVX distance, one heap push) does not depend on which 8-bit codec filled the rows."""
    from oracle import oracle as orc
    st = orc.FlatStore(dim, orc.COSINE, orc.Q_BF16 if wl["quant"] == "bf16" else orc.Q_F8)
    st.upsert(ids, rows)
    return st


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (the port of VertexSearch's control flow
    driving the reference's compiled avx.cpp through oracle/_ref when present), highCpu = 16 shard workers, on all host
    threads it can use.  Honours --steps/--warmup: a step is a bounded sample (q queries x r rows, time scaled linearly in
    rows to the configured shard) sized so that the whole run ends within ~2.5 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    wl = WORKLOADS["flat" if args.workload == "hnsw" else args.workload]
    have_ref = orc.use_reference_kernels(True)
    n_threads = min(16, os.cpu_count() or 1)   # EDGE_MAP_SHARD_COUNT goroutines (edge/constants.go:49)
    n_full = min(args.rows, 1_000_000 if args.dim <= 768 else 250_000)   # host memory / generation time bound
    rows = gen_rows(n_full, args.dim, BASE_SEED)
    ids = shard_ids(n_full, 0)
    queries = gen_rows(args.batch, args.dim, QUERY_SEED)
    st = cpu_store(wl, args.dim, rows, ids)
    sel = orc.NEAREST

    def one(j):
        st.search(queries[j % len(queries)], args.k, high_cpu=True, select_mode=sel, n_threads=n_threads)
    t0 = time.perf_counter(); one(0); t_q = time.perf_counter() - t0           # calibration: one query over n_full rows
    budget = 150.0 / max(1, args.steps + args.warmup)
    q_per_step = int(max(1, min(8, budget // max(t_q, 1e-9))))
    if t_q > budget:        # even one full query is too long for this many steps: shrink the row sample
        n_s = max(10_000, int(n_full * budget / t_q))
        st = cpu_store(wl, args.dim, rows[:n_s], ids[:n_s])
    else:
        n_s = n_full
    for w in range(args.warmup):
        for j in range(q_per_step):
            one(w * q_per_step + j)
    t0 = time.perf_counter()
    for i in range(args.steps):
        for j in range(q_per_step):
            one(i * q_per_step + j)
    t = time.perf_counter() - t0
    # one query over n_s rows costs t/(steps*q); the configured shard costs rows/n_s times that (linear scan)
    qps = (args.steps * q_per_step) / t * (n_s / args.rows)
    sample = f"{q_per_step} queries x {n_s} of {args.rows} rows per step, {args.steps} steps + {args.warmup} warm-up, linear-scan time scaled by rows"
    line = {"impl": "reference", "metric": wl["metric"], "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "quantized rows, f32 math", "data": "synthetic",
            "config": {"workload": wl["name"], "rows_per_gpu": args.rows, "dim": args.dim, "batch": args.batch, "k": args.k, "select": "nearest"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": n_threads, "kind": "port",
                             "kernels": "reference avx.cpp (oracle/_ref)" if have_ref else "scalar lane-order port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- config 3 (HNSW) arm
def latent_rows(n, d, seed, lat=32, chunk=100_000):
    """Vectors with structure a graph index can use: a 32-d Gaussian latent through a fixed random projection + 10 % noise
    (on isotropic N(0,1)^768 data M=16 HNSW itself reaches recall@10 ~ 0.1: the reference's walk, not a bug of either side)."""
    A = np.random.Generator(np.random.Philox(0xA)).standard_normal((lat, d), dtype=np.float32)
    g = np.random.Generator(np.random.Philox(seed))
    out = np.empty((n, d), np.float32)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        out[i:i + m] = g.standard_normal((m, lat), dtype=np.float32) @ A + np.float32(0.1) * g.standard_normal((m, d), dtype=np.float32)
    return out


def hnsw_arm(args, light=False):
    """BASELINE configs[2]: core/vectorindex HNSW fp32 dim=768 N=1M efSearch=128 top-10 on one B200.  A step = one batch
    through coltt_b200_hnsw_search (host buffers in and out: this path has no device-resident entry point, so `value`
    and `e2e` are the same measurement); the roofline is the random row gather of the walk.
    Returns the record (rank 0) — printed by --workload hnsw, embedded compactly (light=True) in the default line."""
    import torch
    import coltt_b200 as cb
    from coltt_b200 import _lib
    from oracle import oracle as orc
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    L = _lib.lib()
    assert L.coltt_b200_device_count() >= 1, "no sm_100 GPU: coltt_b200 has no CPU fallback"
    torch.cuda.set_device(0)
    n, d, k, ef = args.rows, args.dim, args.k, args.ef
    nq = args.batch if args.batch != 256 else 1024
    steps, warmup = min(args.steps, 50), max(3, min(args.warmup, 5))
    rows = latent_rows(n, d, BASE_SEED)
    ids = np.arange(n, dtype=np.uint64) + 1
    t0 = time.perf_counter()
    h = cb.Hnsw.Build(ids, rows, metric=cb.Distance_Cosine, m=16, ef=ef)
    t_build = time.perf_counter() - t0
    qsets = [latent_rows(nq, d, QUERY_SEED + i) for i in range(4)]
    sampler = ClockSampler(0)       # started before the warm-up, rows time-stamped: see flat_arm
    sampler.start()
    t_load0 = time.perf_counter()
    for i in range(warmup):
        h.BatchSearch(qsets[i % 4], k, ef)
    launches0 = L.coltt_b200_kernel_launches()
    kern, evals, exps = [], 0, 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        h.BatchSearch(qsets[i % 4], k, ef)
        st = h.last_stats()
        kern.append(st["kernel_ms"]); evals += st["dist_evals"]; exps += st["expansions"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = L.coltt_b200_kernel_launches() - launches0
    extra = 0
    while time.perf_counter() - t_load0 < 0.45:      # the same searches, untimed, until nvidia-smi has seen the load
        h.BatchSearch(qsets[extra % 4], k, ef)
        extra += 1
    clocks = sampler.stop(t_load0 + 0.05, time.perf_counter())
    clocks["continued_steps_for_sampling"] = extra
    kern_ms = float(np.mean(kern))
    qps = nq * steps / wall
    peaks = measured_peaks()
    alg = (evals * (d * 4 + 8) + exps * 32 * 4) / steps                 # E*(D*4+8) + X*mMax0*4 per batch
    roof = {"bound": "hbm", "achieved": alg / (kern_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": ncu_traffic("hnsw_search_kernel"),
            "traffic_note": "captured at N=100K, batch 256 (profiles/r1_hnsw_search_summary.md)", "kernel": "hnsw_search_kernel",
            "algorithmic_bytes": alg, "kernel_ms": kern_ms, "peak_source": peaks["src"],
            "note": "random 3 KB row gathers; dist_evals and expansions are the oracle's counts"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    line = {"metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * wall / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "core/vectorindex HNSW fp32 cosine dim=768 N=1M M=16 efSearch=128 top-10 (BASELINE configs[2])", "rows": n, "dim": d,
                       "batch": nq, "k": k, "ef": ef, "data_model": "32-d latent + 10% noise", "build_s": round(t_build, 2),
                       "l2": "row table 3.07 GB > 126 MB L2 (inputs larger than L2)", "dist_evals_per_query": evals / (steps * nq),
                       "expansions_per_query": exps / (steps * nq)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 16 + nq * 4 + 64,
                    "note": "value is already end to end (host buffers through the C-ABI)"},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clocks}
    if not args.no_cpu:
        # recall@10 against exact search (FLAT fp32 store on the GPU, bit-identical to the oracle's arithmetic)
        rq = min(64, nq)

        def exact_topk(ids_, rows_, qs_):
            sp = cb.VectorSpace("gt", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None), capacity_hint=len(ids_), select_mode=cb.SELECT_NEAREST)
            sp.ChangedVertices(ids_, rows_)
            wi_, _, _ = sp.BatchVertexSearch(qs_, k, math_mode=cb.MATH_EXACT)
            sp.close()
            return wi_
        wi = exact_topk(ids, rows, qsets[0][:rq])
        gi, _, _ = h.BatchSearch(qsets[0], k, ef)
        line["recall_at_10"] = float(np.mean([orc.compute_recall(wi[j, :k], gi[j, :k], k) for j in range(rq)]))
        # CPU baseline: the oracle's literal hnsw.go walk, one thread (Hnsw.Search is single-threaded per query), on a bounded
        # sample graph of 100 K vertices built here and loaded from its Commit blob (a 1 M-vertex blob is 3.5 GB of Python bytes)
        ns = min(n, 100_000)
        cq = 32

        def sample_graph(rows_s, qs_s, time_it):
            """GPU walk vs the oracle's literal walk on the same 100 K-vertex graph: recall of both against exact search."""
            hs = cb.Hnsw.Build(ids[:ns], rows_s, metric=cb.Distance_Cosine, m=16, ef=ef)
            oh = orc.Hnsw.load(hs.Commit())
            oh.set_ef(ef)
            t0_ = time.perf_counter()
            o_hits = [oh.search(qs_s[j], k)[0] for j in range(cq)]
            t_ = time.perf_counter() - t0_
            g_ids, _, _ = hs.BatchSearch(qs_s[:cq], k, ef)
            hs.close()
            w_ = exact_topk(ids[:ns], rows_s, qs_s[:cq])
            rec_o = float(np.mean([orc.compute_recall(w_[j, :k], o_hits[j], k) for j in range(cq)]))
            rec_g = float(np.mean([orc.compute_recall(w_[j, :k], g_ids[j, :k], k) for j in range(cq)]))
            same = all(np.array_equal(np.asarray(o_hits[j], np.uint64), g_ids[j, :len(o_hits[j])]) for j in range(cq))
            return rec_g, rec_o, same, cq / t_
        rg, ro, same, cpu_qps = sample_graph(rows[:ns], qsets[0], True)
        line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"{cq} queries on a {ns}-vertex graph (of {n}); the walk grows ~log N, so this over-states the CPU at 1 M"}
        line["sample_graph_latent"] = {"vertices": ns, "recall_gpu": rg, "recall_oracle_walk": ro, "gpu_ids_equal_oracle": bool(same)}
        # the same on isotropic N(0,1)^768 data (BASELINE's "synthetic dim-768" read literally): a graph index has nothing to
        # exploit there, and the reference's own walk (the oracle) has the same recall — the latent set is a data-model choice
        iso = gen_rows(ns, d, BASE_SEED + 77)
        iq = gen_rows(cq, d, QUERY_SEED + 77)
        rg, ro, same, _ = sample_graph(iso, iq, False)
        line["isotropic"] = {"vertices": ns, "recall_gpu": rg, "recall_oracle_walk": ro, "gpu_ids_equal_oracle": bool(same),
                             "note": "N(0,1)^768, same M / ef; recall is the reference walk's property, not the port's"}
    h.close()
    return line


# ------------------------------------------------------------------------------------ our arm
def device_rows(torch, dev, n, d, seed, chunk=250_000):
    """Synthetic N(0,1) rows generated on the device chunk by chunk (deterministic per (seed, chunk)): yields (offset, tensor)."""
    for ci, off in enumerate(range(0, n, chunk)):
        g = torch.Generator(device=dev)
        g.manual_seed((seed << 20) + ci)
        yield off, torch.randn((min(chunk, n - off), d), generator=g, device=dev, dtype=torch.float32)


def flat_arm(args, wl, torch, dist, world, rank, local, light=False):
    """One FLAT workload (config 2 or config 4) on this rank's shard; returns the JSON record on rank 0 (None elsewhere).
    light = the compact record embedded in the default line (fewer steps, no e2e / CPU legs)."""
    import coltt_b200 as cb
    from coltt_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", local)
    n, d, nq, k = args.rows, args.dim, args.batch, args.k
    qname = getattr(args, "quant", None) or wl["quant"]
    fp8 = qname == "f8_e4m3"
    quant = {"f8_e4m3": cb.Quantization_F8_E4M3, "bf16": cb.Quantization_BF16, "none": cb.Quantization_None}[qname]
    math_mode = {"auto": cb.MATH_FAST,   # tcgen05 filter + exact re-rank: bit-identical results to EXACT (tests/test_gpu_fast.py, test_gpu_f8e.py)
                 "exact": cb.MATH_EXACT, "fast": cb.MATH_FAST}[args.math]
    t0 = time.perf_counter()
    sp = cb.VectorSpace("bench", cb.Metadata(d, cb.Distance_Cosine, quant), device=local, capacity_hint=n,
                        select_mode=cb.SELECT_NEAREST, math_mode=math_mode)
    id_base = (rank + 1) << 40
    rows = ids = None
    if wl["device_gen"]:
        for off, t in device_rows(torch, dev, n, d, BASE_SEED + rank):
            sp.AppendDeviceRows(t.data_ptr(), t.shape[0], d, id_base)
        torch.cuda.synchronize(dev)
    else:
        rows = gen_rows(n, d, BASE_SEED + rank)
        ids = shard_ids(n, rank)
        sp.ChangedVertices(ids, rows)
    t_ingest = time.perf_counter() - t0

    n_qsets = 4  # distinct query batches cycled across steps
    q_host = [gen_rows(nq, d, QUERY_SEED + i) for i in range(n_qsets)]
    q_dev = [torch.from_numpy(q).to(dev) for q in q_host]
    out = torch.zeros((nq, k, 4), dtype=torch.int32, device=dev)      # coltt_hit = 16 B
    cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    comm = None
    if world > 1:
        # the sharded search lives behind the C-ABI (csrc/comm.cu): local search -> ONE ncclAllGather of per-shard top-k ->
        # K5 merge, on one stream inside the library; torch.distributed only carries the rendezvous blob and the barriers
        from coltt_b200.dist import Comm
        comm = Comm.from_torch_distributed(local)
        fin = torch.zeros((nq, k, 4), dtype=torch.int32, device=dev)
        fcnt = torch.zeros((nq,), dtype=torch.int32, device=dev)

    def step_dev(i):
        if world > 1:
            comm.search_dev(sp, q_dev[i % n_qsets].data_ptr(), nq, k, cb.SELECT_NEAREST, math_mode, fin.data_ptr(), fcnt.data_ptr(), stream.cuda_stream)
        else:
            _lib.check(L.coltt_b200_store_search_dev(sp._h, q_dev[i % n_qsets].data_ptr(), nq, k, cb.SELECT_NEAREST, math_mode,
                                                      out.data_ptr(), cnt.data_ptr(), stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    steps, warmup = (min(args.steps, 10), 3) if light else (args.steps, args.warmup)
    # ---- device-resident throughput -------------------------------------------------------
    # nvidia-smi needs ~0.2 s to print its first row, longer than a 20-step timed region (8 ms at config 2): the sampler starts
    # before the warm-up, and when warm-up + timed region lasted less than 0.45 s the SAME steps keep running, untimed, right
    # behind the timed region until 0.45 s of load have passed.  Rows are time-stamped; only rows printed while the device was
    # under this load (warm-up, timed region, continuation) are summarised.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load0 = time.perf_counter()
    for i in range(warmup):
        step_dev(i)
    barrier()
    launches0 = L.coltt_b200_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    launches = L.coltt_b200_kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # the continuation: every rank runs the same number of extra steps (the step is a collective at N > 1)
    load_ms = (warmup + steps) * ms / steps
    extra = int(min(5000, max(0.0, 450.0 - load_ms) / (ms / steps) + 0.999)) if load_ms < 450.0 else 0
    for i in range(extra):
        step_dev(i)
        if i % 64 == 63:
            torch.cuda.synchronize(dev)
    barrier()
    clocks = None
    if rank == 0:
        clocks = sampler.stop(t_load0 + 0.05, time.perf_counter())
        clocks["continued_steps_for_sampling"] = extra
        if extra:
            clocks["note"] = (f"timed region {ms:.1f} ms is below nvidia-smi's resolution: rows cover warm-up + timed region + {extra} more "
                              "untimed steps of the same load (0.45 s in all), long enough for the power cap to engage")
    ms_per_step = ms / steps
    value = world * nq * steps / (ms / 1000.0)   # (query x shard) units per second, whole job

    # ---- dominant-kernel time, per launch, with CUDA events on its stream (instrumented pass)
    scan_ms = []
    sp.set_timing(True)          # per-phase events are off during the timed regions above and below
    for i in range(min(steps, 20)):
        step_dev(i)
        torch.cuda.synchronize(dev)
        scan_ms.append(sp.last_timing_ms())
    sp.set_timing(False)
    kern_ms = float(np.mean([x["scan"] for x in scan_ms]))
    parts = {kk: float(np.mean([x[kk] for x in scan_ms])) for kk in ("prep", "scan", "rerank", "merge")}

    # ---- merged-result check at N>1: rank 0 recomputes sampled queries over the UNION of the shards from per-rank
    # EXACT searches merged on the host (no NCCL, no device merge on that path) and compares ids + score bits
    merge_check = None
    if world > 1:
        mq = min(8, nq)
        step_dev(0)
        torch.cuda.synchronize(dev)
        got = fin[:mq].cpu().numpy().view(np.uint8).reshape(mq, k, 16)
        gi_, gs_ = got[..., :8].copy().view(np.uint64)[..., 0], got[..., 8:12].copy().view(np.float32)[..., 0]
        ei, es, ec = sp.BatchVertexSearch(q_host[0][:mq], k, select_mode=cb.SELECT_NEAREST, math_mode=cb.MATH_EXACT)
        loc = torch.from_numpy(np.concatenate([ei.view(np.int64).astype(np.float64), es.astype(np.float64)], axis=1)).to(dev)
        allp = [torch.zeros_like(loc) for _ in range(world)] if rank == 0 else None
        dist.gather(loc, allp, dst=0)
        if rank == 0:
            ok = True
            for j in range(mq):
                cid = np.concatenate([a[j, :k].cpu().numpy().astype(np.int64).view(np.uint64) for a in allp])
                csc = np.concatenate([a[j, k:].cpu().numpy().astype(np.float32) for a in allp])
                order = np.lexsort((cid, csc))[:k]
                ok &= bool(np.array_equal(cid[order], gi_[j]) and csc[order].tobytes() == gs_[j].tobytes())
            merge_check = "ok" if ok else "MISMATCH"

    # ---- end to end: host buffers in, host results out.  N=1: the host-pointer C-ABI call.  N>1: the sharded
    # C-ABI call coltt_b200_sharded_search: H2D of the queries, per-shard search, all-gather, merge, D2H.
    e2e_value = e2e_steps = e2e_pageable = None
    if not light:
        # the caller's query batches live in page-locked buffers (coltt_b200_host_alloc), as the bench contract asks: the
        # library DMAs them in place.  The same loop over ordinary (pageable) numpy arrays — staged through the handle's pinned
        # buffer by one extra host memcpy — is reported beside it.
        q_pin = []
        for q in q_host:
            b = cb.pinned_empty(q.shape, np.float32)
            b[...] = q
            q_pin.append(b)

        def e2e_leg(qsets):
            if world > 1:
                def e2e_step(i):     # coltt_b200_sharded_search: host queries in, merged host results out, on every rank
                    return comm.search(sp, qsets[i % n_qsets], k, cb.SELECT_NEAREST, math_mode)
            else:
                def e2e_step(i):
                    return sp.BatchVertexSearch(qsets[i % n_qsets], k)
            for i in range(max(3, min(warmup, 10))):
                e2e_step(i)
            barrier()
            t1 = time.perf_counter()
            n_e = max(50, min(steps // 3, 300)) if nq * d <= 256 * 768 else max(5, min(steps // 3, 300))
            for i in range(n_e):
                e2e_step(i)
            torch.cuda.synchronize(dev)
            e2e_s = time.perf_counter() - t1
            if dist is not None:
                t = torch.tensor([e2e_s], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            return world * nq * n_e / e2e_s, n_e
        e2e_pageable, _ = e2e_leg(q_host)
        e2e_value, e2e_steps = e2e_leg(q_pin)

    # ---- recall@k against fp32 ground truth.  Host-resident rows: the oracle.  Device-generated shards (config 4): the
    # rows are regenerated chunk by chunk and scored in fp32 by torch (CUDA-core matmul, TF32 off) — a checker independent
    # of this library's kernels; the oracle covers the same store at 100 K rows in tests/test_gpu_f8e.py.
    recall = recall_note = None
    if not args.no_cpu and world == 1 and wl["device_gen"]:
        rq = min(args.recall_queries, nq)
        torch.backends.cuda.matmul.allow_tf32 = False
        qn = torch.nn.functional.normalize(q_dev[0][:rq], dim=1)
        best_s = torch.full((rq, k), -2.0, device=dev)
        best_i = torch.zeros((rq, k), dtype=torch.int64, device=dev)
        for off, t in device_rows(torch, dev, n, d, BASE_SEED + rank):
            sim = qn @ torch.nn.functional.normalize(t, dim=1).T
            s_, i_ = sim.topk(min(k, sim.shape[1]), dim=1)
            cs, ci = torch.cat([best_s, s_], 1), torch.cat([best_i, i_ + off + id_base], 1)
            best_s, sel_ = cs.topk(k, dim=1)
            best_i = ci.gather(1, sel_)
        gi, _, _ = sp.BatchVertexSearch(q_host[0][:rq], k)
        gt = best_i.cpu().numpy().astype(np.uint64)
        recall = float(np.mean([len(np.intersect1d(gt[j], gi[j, :k])) / k for j in range(rq)]))   # edge/resultset.go:55-65
        recall_note = f"{rq} queries vs fp32 exact ground truth (torch fp32 matmul over the regenerated rows) over all {n} rows"

    exchange = comm.exchange if comm is not None else None
    if comm is not None:
        barrier()           # no peer is still reading this rank's exchange buffer (coltt_b200_comm_destroy's contract)
        comm.close()
    if rank != 0:
        sp.close()
        return None

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = measured_peaks()
    es = {"f8_e4m3": 1, "bf16": 2, "none": 2}[qname]      # bytes per element the dominant kernel streams (fp32 stores: the fp16 shadow)
    passes = (nq + 7) // 8 if math_mode == cb.MATH_EXACT else 1
    alg_bytes = n * d * es + n * 4 + (n * 4 if fp8 else 0) + nq * d * 4 + nq * k * 16          # SURVEY §8(d), one pass (+ row scales)
    flops = 2.0 * nq * n * d
    if math_mode == cb.MATH_EXACT:
        # exact path: CUDA-core fp32 (unfused mul+add): bound by the FP32 pipe, reported against HBM for the
        # bytes it must move (one pass per 8 queries) and as fp32 FLOP/s
        roof = {"bound": "hbm", "achieved": passes * alg_bytes / (kern_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "traffic": None, "note": f"exact-order CUDA-core path: {passes} HBM passes (8 queries each); "
                                         f"{flops / (kern_ms / 1e3) / 1e12:.1f} TFLOP/s fp32 unfused"}
    else:
        kname = "gemm_filter_pair_kernel" if nq > 128 else "gemm_filter_kernel"
        roof = roofline_fast(alg_bytes, flops, kern_ms, ms_per_step, peaks, ncu_traffic(kname + ("_fp8" if fp8 else "")), kname, fp8)
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = peaks["src"]
    roof["kernel_ms"] = kern_ms

    unit = "queries/s" if world == 1 else "shard-queries/s"
    line = {"metric": wl["metric"], "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["name"] if qname == wl["quant"] else wl["name"] + f" [store quantization overridden: {qname}]",
                       "rows_per_gpu": n, "dim": d, "batch": nq, "k": k, "select": "nearest",
                       "math": "exact" if math_mode == cb.MATH_EXACT else "fast(tcgen05)+exact rerank",
                       "l2": f"shard {n * d * es / 1e9:.2f} GB > 126 MB L2 (inputs larger than L2)", "unit_note":
                       "weak scaling: every rank holds its own shard of rows_per_gpu rows and every query is answered over all "
                       "of them; value counts (query x shard) units = n_gpus x batch per step; global_qps is queries/s over the "
                       "whole n_gpus x rows_per_gpu collection",
                       "parallelism": f"shard{world}", "exchange": exchange, "ingest_s": round(t_ingest, 2)},
            "global_qps": nq * steps / (ms / 1000.0), "global_rows": world * n,
            "gpu_launches": int(launches), "kernel_ms": parts, "roofline": roof, "clocks": clocks}
    if e2e_value is not None:
        line["e2e"] = {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 16 + nq * 4,
                       "steps": e2e_steps, "global_qps": e2e_value / world, "host_buffers": "page-locked (coltt_b200_host_alloc), DMA in place",
                       "value_pageable_host_buffers": e2e_pageable}
    if merge_check is not None:
        line["merge_check"] = merge_check
        line["merge_check_note"] = "8 queries: device all-gather + K5 merge vs host merge of per-rank EXACT searches over the union, ids + score bits"
    if math_mode == cb.MATH_FAST:
        fs = sp.fast_stats()
        # queries answered by the tensor-core filter, and how many of them the certificate sent to the exact re-run
        line["fast_path"] = {"queries": fs["queries"], "exact_reruns": fs["exact_reruns"],
                             "rerun_rate": (fs["exact_reruns"] / fs["queries"]) if fs["queries"] else None}
    if recall is not None:
        line[f"recall_at_{k}"] = recall
        line["recall_note"] = recall_note

    if not args.no_cpu and world == 1 and not light:
        from oracle import oracle as orc
        n_threads = min(16, os.cpu_count() or 1)
        have_ref = orc.use_reference_kernels(True)
        n_s = min(args.cpu_rows if d <= 768 else args.cpu_rows // 4, n)
        c_rows = rows[:n_s] if rows is not None else gen_rows(n_s, d, BASE_SEED)
        c_ids = ids[:n_s] if ids is not None else shard_ids(n_s, 0)
        st = cpu_store(wl, d, c_rows, c_ids)
        cq = min(args.cpu_queries, nq)
        t0 = time.perf_counter()
        for j in range(cq):
            st.search(q_host[0][j], k, high_cpu=True, select_mode=orc.NEAREST, n_threads=n_threads)  # highCpu: 16 shard workers
        t = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cq / t * (n_s / n), "unit": "queries/s", "cores": n_threads, "kind": "port",
                                "kernels": "reference avx.cpp (oracle/_ref)" if have_ref else "scalar lane-order port",
                                "sample": f"{cq} queries x {n_s} of {n} rows, highCpu (16 shard workers), scaled by rows; "
                                          "contiguous rows, no Go map/alloc/GC => faster than the real Go path"}
        if rows is not None:
            # recall@10 (edge/resultset.go:55-65) of the fp16 store vs fp32 ground truth from the oracle
            gt = orc.FlatStore(d, orc.COSINE, orc.Q_NONE)
            gt.upsert(ids, rows)
            rq = min(args.recall_queries, nq)
            gi, gs, gc = sp.BatchVertexSearch(q_host[0][:rq], k)
            rec, exact_ok = [], True
            own = orc.FlatStore(d, orc.COSINE, orc.Q_BF16)     # the oracle over the SAME store: ids + score bits at full size
            own.upsert(ids, rows)
            for j in range(rq):
                wi, _ = gt.search_total_order(q_host[0][j], k, select_mode=orc.NEAREST, n_threads=os.cpu_count() or 1)
                rec.append(orc.compute_recall(wi, gi[j, :k], k))
                oi, os_ = own.search_total_order(q_host[0][j], k, select_mode=orc.NEAREST, n_threads=os.cpu_count() or 1)
                exact_ok &= bool(np.array_equal(oi, gi[j, :gc[j]]) and os_.tobytes() == gs[j, :gc[j]].tobytes())
            line[f"recall_at_{k}"] = float(np.mean(rec))
            line["recall_note"] = f"{rq} queries vs fp32 exact ground truth (oracle) over all {n} rows"
            line["oracle_parity_at_full_size"] = "ok" if exact_ok else "MISMATCH"
    sp.close()
    return line


def c5_arm(args, wl, torch, dist, world, rank, local):
    """BASELINE configs[4]: HNSW + PQ, dim 768, efSearch 256, one independent sub-graph per GPU over its row shard,
    product-quantized walk (csrc/pq.cu) + exact re-rank of the ef survivors, ONE all-gather of per-shard top-k + merge through
    the C-ABI communicator.  PARITY UNPINNED (the reference has no PQ arithmetic): judged on recall@10 against exact search
    over the union of the shards.  A step = one batch through coltt_b200_sharded_hnsw_pq_search (host buffers)."""
    import coltt_b200 as cb
    from coltt_b200 import _lib
    from coltt_b200.dist import Comm
    L = _lib.lib()
    dev = torch.device("cuda", local)
    n, d, k, ef, nq = args.rows, args.dim, args.k, args.ef, args.batch
    t0 = time.perf_counter()
    rows = latent_rows(n, d, BASE_SEED + rank)
    ids = (np.arange(n, dtype=np.uint64) + np.uint64(1)) + (np.uint64(rank) << np.uint64(40))
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = cb.Hnsw.Build(ids, rows, metric=cb.Distance_Cosine, m=16, ef=ef, device=local)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    h.TrainPQ(num_centroids=256, num_sub_vectors=64, trigger_threshold=65536)
    t_pq = time.perf_counter() - t0
    comm = Comm.from_torch_distributed(local)
    qsets = [latent_rows(nq, d, QUERY_SEED + i) for i in range(4)]

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()

    for i in range(args.warmup):
        comm.hnsw_search(h, qsets[i % 4], k, ef, pq=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.coltt_b200_kernel_launches()
    kern, evals, exps = [], 0, 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        gi, gs, gc = comm.hnsw_search(h, qsets[i % 4], k, ef, pq=True)
        st = h.last_stats()
        kern.append(st["kernel_ms"]); evals += st["dist_evals"]; exps += st["expansions"]
    barrier()
    wall = time.perf_counter() - t0
    launches = L.coltt_b200_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([wall], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    # ---- recall@10 against exact search over the union: per-shard exact fp32 top-k, merged on the host of rank 0
    rq = min(64, nq)
    gi, gs, gc = comm.hnsw_search(h, qsets[0], k, ef, pq=True)
    fi, _, _ = comm.hnsw_search(h, qsets[0], k, ef, pq=False)            # the fp32 walk over the same sub-graphs, for comparison
    sp = cb.VectorSpace("gt", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None), device=local, capacity_hint=n, select_mode=cb.SELECT_NEAREST)
    sp.ChangedVertices(ids, rows)
    wi, ws, _ = sp.BatchVertexSearch(qsets[0][:rq], k, math_mode=cb.MATH_EXACT)
    sp.close()
    loc = torch.from_numpy(np.concatenate([wi.view(np.int64).astype(np.float64), ws.astype(np.float64)], axis=1)).to(dev)
    if dist is not None:
        allp = [torch.zeros_like(loc) for _ in range(world)] if rank == 0 else None
        dist.gather(loc, allp, dst=0)
    else:
        allp = [loc]
    exchange = comm.exchange if comm is not None else None
    if comm is not None:
        barrier()
        comm.close()
    if rank != 0:
        h.close()
        return None
    rec_pq, rec_f32 = [], []
    for j in range(rq):
        cid = np.concatenate([a[j, :k].cpu().numpy().astype(np.int64).view(np.uint64) for a in allp])
        csc = np.concatenate([a[j, k:].cpu().numpy() for a in allp])
        truth = cid[np.lexsort((cid, csc))[:k]]
        rec_pq.append(len(np.intersect1d(truth, gi[j, :k])) / k)
        rec_f32.append(len(np.intersect1d(truth, fi[j, :k])) / k)
    kern_ms = float(np.mean(kern))
    peaks = measured_peaks()
    M = 64
    alg = (evals * M + exps * 32 * 4 + args.steps * nq * ef * (d * 4 + 8)) / args.steps        # codes + lists + re-ranked rows, per batch (this rank)
    roof = {"bound": "hbm", "achieved": alg / (kern_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None,
            "kernel": "pq_search_kernel + pq_finish_kernel", "algorithmic_bytes": alg, "kernel_ms": kern_ms, "peak_source": peaks["src"],
            "note": "latency-bound graph walk: E x 64 B code gathers + X x 128 B lists + ef x 3 KB re-ranked rows per query (SURVEY 8d: bytes = E x M)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    qps = nq * args.steps / wall
    unit = "queries/s" if world == 1 else "shard-queries/s"
    line = {"metric": wl["metric"], "value": world * qps, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
            "data": "synthetic",
            "config": {"workload": wl["name"], "rows_per_gpu": n, "dim": d, "batch": nq, "k": k, "ef": ef, "pq": "64 sub-vectors x 256 centroids, 65536 training rows",
                       "data_model": "32-d latent + 10% noise", "gen_s": round(t_gen, 1), "build_s": round(t_build, 1), "pq_train_encode_s": round(t_pq, 2),
                       "parallelism": f"shard{world}", "exchange": exchange, "code_evals_per_query": evals / (args.steps * nq), "expansions_per_query": exps / (args.steps * nq),
                       "unit_note": "weak scaling (rows_per_gpu per rank); value counts (query x shard) units, global_qps is queries/s over the union"},
            "global_qps": qps, "global_rows": world * n,
            "e2e": {"value": world * qps, "unit": unit, "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 16 + nq * 4,
                    "note": "value is already end to end (host buffers through the C-ABI)", "global_qps": qps},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
            "recall_at_10": float(np.mean(rec_pq)), "recall_at_10_fp32_walk": float(np.mean(rec_f32)),
            "recall_note": f"{rq} queries vs exact fp32 search over the union of the {world} shards; parity unpinned (builder-defined PQ, SURVEY F5)"}
    h.close()
    return line


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "hnsw":
        rec = hnsw_arm(args)
        if rec is not None:
            print(json.dumps(rec), flush=True)
        return

    import torch
    from coltt_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(torch.device("cuda", local))
    assert _lib.lib().coltt_b200_device_count() >= 1, "no sm_100 GPU: coltt_b200 has no CPU fallback"

    wl = WORKLOADS[args.workload]
    if args.workload == "c5":
        line = c5_arm(args, wl, torch, dist, world, rank, local)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return
    line = flat_arm(args, wl, torch, dist, world, rank, local)
    # The default line also carries compact records of configs[2] (HNSW) and configs[3] (fp8, top-100, batch 1024) on this
    # GPU, so that they are driver-run numbers; time-boxed, N=1 only.
    if rank == 0 and world == 1 and args.workload == "flat" and not args.no_extras and not args.no_cpu:
        try:
            a4 = argparse.Namespace(**vars(args))
            w4 = WORKLOADS["c4"]
            for key in ("rows", "dim", "batch", "k", "steps", "warmup", "quant"):
                setattr(a4, key, w4[key])
            r4 = flat_arm(a4, w4, torch, None, 1, 0, local, light=True)
            line["c4"] = {kk: r4[kk] for kk in ("value", "unit", "ms_per_step", "steps", "kernel_ms", "fast_path", "clocks", "recall_at_100", "recall_note") if kk in r4}
            line["c4"]["config"] = r4["config"]["workload"]
            line["c4"]["roofline"] = {kk: r4["roofline"][kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "tensor_peak_note", "hbm_frac", "kernel_ms")}
            line["c4"]["parity"] = "reference parity unpinned (builder-defined E4M3 store); FAST == EXACT == oracle restatement in tests/test_gpu_f8e.py"
        except Exception as e:      # the headline line must survive a failure of an extra leg
            line["c4"] = {"error": repr(e)[:300]}
        try:
            a3 = argparse.Namespace(**vars(args))
            a3.rows, a3.dim, a3.batch, a3.k, a3.steps, a3.warmup = 1_000_000, 768, 1024, 10, 10, 3
            r3 = hnsw_arm(a3, light=True)
            line["c3"] = {kk: r3[kk] for kk in ("value", "unit", "ms_per_step", "steps", "recall_at_10", "isotropic", "clocks") if kk in r3}
            line["c3"]["config"] = r3["config"]["workload"]
            line["c3"]["roofline"] = {kk: r3["roofline"][kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms")}
        except Exception as e:
            line["c3"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
