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
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flat", choices=["flat", "hnsw"],
                    help="flat = BASELINE configs[1] (the headline, default); hnsw = configs[2] (core/vectorindex HNSW, one GPU)")
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--math", default="auto", choices=["auto", "exact", "fast"])
    ap.add_argument("--cpu-rows", type=int, default=250_000, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-queries", type=int, default=8)
    ap.add_argument("--recall-queries", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / recall legs (profiling runs)")
    return ap.parse_args()


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
            self.rows.append(line.strip())

    def stop(self):
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
        for r in self.rows:
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
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j.get("hbm_gbs", 6650.0), "tf": j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0)), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf": 1590.0, "src": "fallback"}


def roofline_fast(alg_bytes, flops, kern_ms, peaks, traffic, kname):
    """Roofline of the tcgen05 filter kernel.  Both floors are computed from the MEASURED peaks; the larger one binds:
    at 256 queries x fp16 rows the intensity is 256 flop/B, above the measured ridge (bf16 TFLOP/s / HBM GB/s ~ 210),
    so the tensor pipe is the bound and HBM the secondary figure (both are reported)."""
    sec = kern_ms / 1e3
    gbs, tfs = alg_bytes / sec / 1e9, flops / sec / 1e12
    t_hbm, t_tensor = alg_bytes / (peaks["hbm_gbs"] * 1e9), flops / (peaks["tf"] * 1e12)
    roof = {"kernel": kname, "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops, "traffic": traffic,
            "hbm_gbs": gbs, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_frac": gbs / peaks["hbm_gbs"],
            "tensor_tflops": tfs, "tensor_peak_tflops": peaks["tf"], "tensor_frac": tfs / peaks["tf"],
            "floor_ms": {"hbm": t_hbm * 1e3, "tensor": t_tensor * 1e3}}
    if t_tensor >= t_hbm:
        roof.update({"bound": "tensor", "achieved": tfs, "peak": peaks["tf"], "unit": "TFLOP/s"})
    else:
        roof.update({"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s"})
    return roof


# ------------------------------------------------------------------------------- reference arm
def cpu_arm(args, rows, ids, queries, n_threads):
    """The reference's CPU path (port of VertexSearch control flow driving the reference's own
    avx.cpp through oracle/_ref when present) on a bounded sample; QPS scaled to `args.rows`."""
    from oracle import oracle as orc
    kind = "port"
    have_ref = orc.use_reference_kernels(True)
    n_s = min(args.cpu_rows, rows.shape[0])
    st = orc.FlatStore(args.dim, orc.COSINE, orc.Q_BF16)
    st.upsert(ids[:n_s], rows[:n_s])
    nq = min(args.cpu_queries, queries.shape[0])

    def run():
        t0 = time.perf_counter()
        for j in range(nq):
            st.search(queries[j], args.k, high_cpu=True, select_mode=orc.NEAREST, n_threads=n_threads)  # highCpu: 16 shard workers
        return time.perf_counter() - t0
    return st, n_s, nq, run, kind, have_ref


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = min(16, os.cpu_count() or 1)   # EDGE_MAP_SHARD_COUNT goroutines (edge/constants.go:49)
    rows = gen_rows(min(args.cpu_rows, args.rows), args.dim, BASE_SEED)
    ids = shard_ids(rows.shape[0], 0)
    queries = gen_rows(args.batch, args.dim, QUERY_SEED)
    st, n_s, nq, run, kind, have_ref = cpu_arm(args, rows, ids, queries, n_threads)
    for _ in range(min(args.warmup, 1)):
        run()
    steps = max(1, min(args.steps, 5))
    t = sum(run() for _ in range(steps))
    # one query over n_s rows costs t/(steps*nq); a 1M-row shard costs rows/n_s times that
    qps = (steps * nq) / t * (n_s / args.rows)
    sample = f"{nq} queries x {n_s} of {args.rows} rows per step, {steps} steps, linear-scan time scaled by rows"
    line = {"impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 rows, f32 math", "data": "synthetic",
            "config": {"workload": "edge FLAT bf16(=fp16) cosine dim=768 N=1M batch=256 top-10", "rows": args.rows, "dim": args.dim,
                       "k": args.k, "select": "nearest"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": n_threads, "kind": kind,
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


def run_hnsw_arm(args):
    """BASELINE configs[2]: core/vectorindex HNSW fp32 dim=768 N=1M efSearch=128 top-10 on one B200.  A step = one batch
    through coltt_b200_hnsw_search (host buffers in and out: this path has no device-resident entry point, so `value`
    and `e2e` are the same measurement); the roofline is the random row gather of the walk."""
    import torch
    import coltt_b200 as cb
    from coltt_b200 import _lib
    if int(os.environ.get("RANK", "0")) != 0:
        return
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
    for i in range(warmup):
        h.BatchSearch(qsets[i % 4], k, ef)
    sampler = ClockSampler(0)
    sampler.start()
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
    clocks = sampler.stop()
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
        from oracle import oracle as orc
        # recall@10 against exact search (FLAT fp32 store on the GPU, bit-identical to the oracle's arithmetic)
        rq = min(64, nq)
        sp = cb.VectorSpace("gt", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None), capacity_hint=n, select_mode=cb.SELECT_NEAREST)
        sp.ChangedVertices(ids, rows)
        wi, _, _ = sp.BatchVertexSearch(qsets[0][:rq], k, math_mode=cb.MATH_EXACT)
        sp.close()
        gi, _, _ = h.BatchSearch(qsets[0], k, ef)
        line["recall_at_10"] = float(np.mean([orc.compute_recall(wi[j, :k], gi[j, :k], k) for j in range(rq)]))
        # CPU baseline: the oracle's literal hnsw.go walk, one thread (Hnsw.Search is single-threaded per query), on a bounded
        # sample graph of 100 K vertices built here and loaded from its Commit blob (a 1 M-vertex blob is 3.5 GB of Python bytes)
        ns = min(n, 100_000)
        hs = cb.Hnsw.Build(ids[:ns], rows[:ns], metric=cb.Distance_Cosine, m=16, ef=ef)
        oh = orc.Hnsw.load(hs.Commit())
        hs.close()
        oh.set_ef(ef)
        cq = 32
        t0 = time.perf_counter()
        for j in range(cq):
            oh.search(qsets[0][j], k)
        line["cpu_baseline"] = {"value": cq / (time.perf_counter() - t0), "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"{cq} queries on a {ns}-vertex graph (of {n}); the walk grows ~log N, so this over-states the CPU at 1 M"}
    print(json.dumps(line), flush=True)
    h.close()


# ------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "hnsw":
        return run_hnsw_arm(args)

    import torch
    import coltt_b200 as cb
    from coltt_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    assert L.coltt_b200_device_count() >= 1, "no sm_100 GPU: coltt_b200 has no CPU fallback"

    n, d, nq, k = args.rows, args.dim, args.batch, args.k
    math_mode = {"auto": cb.MATH_FAST,   # tcgen05 filter + exact re-rank: bit-identical results to EXACT (tests/test_gpu_fast.py)
                 "exact": cb.MATH_EXACT, "fast": cb.MATH_FAST}[args.math]
    t0 = time.perf_counter()
    rows = gen_rows(n, d, BASE_SEED + rank)
    ids = shard_ids(n, rank)
    sp = cb.VectorSpace("bench", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_BF16), device=local, capacity_hint=n,
                        select_mode=cb.SELECT_NEAREST, math_mode=math_mode)
    sp.ChangedVertices(ids, rows)
    t_ingest = time.perf_counter() - t0

    n_qsets = 4  # distinct query batches cycled across steps
    q_host = [gen_rows(nq, d, QUERY_SEED + i) for i in range(n_qsets)]
    q_dev = [torch.from_numpy(q).to(dev) for q in q_host]
    out = torch.zeros((nq, k, 4), dtype=torch.int32, device=dev)      # coltt_hit = 16 B
    cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    if world > 1:
        packed = torch.zeros((nq * k * 4 + nq,), dtype=torch.int32, device=dev)          # hits + counts of this shard
        gathered_flat = torch.zeros((world * (nq * k * 4 + nq),), dtype=torch.int32, device=dev)
        fin = torch.zeros((nq, k, 4), dtype=torch.int32, device=dev)
        fcnt = torch.zeros((nq,), dtype=torch.int32, device=dev)

    brk = os.environ.get("COLTT_BENCH_BREAKDOWN") == "1"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if brk else None
    brk_parts = {"search": 0.0, "gather": 0.0, "merge": 0.0, "n": 0}

    def step_dev(i):
        if brk:
            ev[0].record(stream)
        _lib.check(L.coltt_b200_store_search_dev(sp._h, q_dev[i % n_qsets].data_ptr(), nq, k, cb.SELECT_NEAREST, math_mode,
                                                  out.data_ptr(), cnt.data_ptr(), stream.cuda_stream))
        if world > 1:
            # the one exchange step of the sharded search: all-gather of per-shard top-k (counts ride along in
            # the same message), then the K5 merge on every rank
            with torch.cuda.stream(stream):
                if brk:
                    ev[1].record(stream)
                packed[:nq * k * 4].copy_(out.view(-1), non_blocking=True)
                packed[nq * k * 4:].copy_(cnt, non_blocking=True)
                dist.all_gather_into_tensor(gathered_flat, packed)
                if brk:
                    ev[2].record(stream)
            _lib.check(L.coltt_b200_merge_topk_dev2(local, gathered_flat.data_ptr(), world, nq, k, k, cb.SELECT_NEAREST, (nq * k * 4 + nq) * 4,
                                                     nq * k * 16, fin.data_ptr(), fcnt.data_ptr(), stream.cuda_stream))
            if brk:
                ev[3].record(stream)
                torch.cuda.synchronize(dev)
                brk_parts["search"] += ev[0].elapsed_time(ev[1])
                brk_parts["gather"] += ev[1].elapsed_time(ev[2])
                brk_parts["merge"] += ev[2].elapsed_time(ev[3])
                brk_parts["n"] += 1

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -------------------------------------------------------
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.coltt_b200_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    launches = L.coltt_b200_kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * nq * args.steps / (ms / 1000.0)   # shard-queries/s, whole job

    # ---- dominant-kernel time, per launch, with CUDA events on its stream (instrumented pass)
    scan_ms = []
    sp.set_timing(True)          # per-phase events are off during the timed regions above and below
    for i in range(min(args.steps, 20)):
        step_dev(i)
        torch.cuda.synchronize(dev)
        scan_ms.append(sp.last_timing_ms())
    sp.set_timing(False)
    kern_ms = float(np.mean([x["scan"] for x in scan_ms]))
    parts = {kk: float(np.mean([x[kk] for x in scan_ms])) for kk in ("prep", "scan", "rerank", "merge")}

    # ---- end to end: host buffers in, host results out.  N=1: the host-pointer C-ABI call.  N>1: the sharded
    # public API (coltt_b200.dist.ShardedSearch): H2D of the queries, per-shard search, all-gather, merge, D2H.
    if world > 1:
        from coltt_b200.dist import ShardedSearch, cuda_callables, unpack_hits
        ls, mg = cuda_callables(sp, local, math_mode=math_mode)
        sharded = ShardedSearch(ls, mg)

        def e2e_step(i):
            hits, c2 = sharded.search(q_host[i % n_qsets], k, cb.SELECT_NEAREST)
            return unpack_hits(hits, c2)
    else:
        def e2e_step(i):
            return sp.BatchVertexSearch(q_host[i % n_qsets], k)
    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(3, min(args.steps // 3, 300))
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t1
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * nq * e2e_steps / e2e_s
    if brk and brk_parts["n"] and rank == 0:
        print("[bench breakdown ms/step]", {kk: round(v / brk_parts["n"], 4) for kk, v in brk_parts.items() if kk != "n"}, file=sys.stderr, flush=True)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = measured_peaks()
    passes = (nq + 7) // 8 if math_mode == cb.MATH_EXACT else 1
    alg_bytes = n * d * 2 + n * 4 + nq * d * 4 + nq * k * 16          # SURVEY §8(d), one pass
    flops = 2.0 * nq * n * d
    if math_mode == cb.MATH_EXACT:
        # exact path: CUDA-core fp32 (unfused mul+add): bound by the FP32 pipe, reported against HBM for the
        # bytes it must move (one pass per 8 queries) and as fp32 FLOP/s
        roof = {"bound": "hbm", "achieved": passes * alg_bytes / (kern_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "traffic": None, "note": f"exact-order CUDA-core path: {passes} HBM passes (8 queries each); "
                                         f"{flops / (kern_ms / 1e3) / 1e12:.1f} TFLOP/s fp32 unfused"}
    else:
        kname = "gemm_filter_pair_kernel" if nq > 128 else "gemm_filter_kernel"
        roof = roofline_fast(alg_bytes, flops, kern_ms, peaks, ncu_traffic(kname), kname)
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = peaks["src"]
    roof["kernel_ms"] = kern_ms

    line = {"metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 rows/queries, f32 accumulate", "data": "synthetic",
            "config": {"workload": "edge FLAT bf16(=fp16) cosine dim=768 N=1M/GPU batch=256 top-10 (BASELINE configs[1])",
                       "rows_per_gpu": n, "dim": d, "batch": nq, "k": k, "select": "nearest",
                       "math": "exact" if math_mode == cb.MATH_EXACT else "fast(tcgen05)+exact rerank",
                       "l2": "shard 1.54 GB > 126 MB L2 (inputs larger than L2)", "unit_note":
                       "value counts each query once per 1M-row shard it is answered over (n_gpus x batch per step)",
                       "parallelism": f"shard{world}", "ingest_s": round(t_ingest, 2)},
            "global_qps": nq * args.steps / (ms / 1000.0),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 16 + nq * 4,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "kernel_ms": parts, "roofline": roof, "clocks": clocks}
    if math_mode == cb.MATH_FAST:
        fs = (C.c_uint64 * 2)()
        _lib.check(L.coltt_b200_store_fast_stats(sp._h, fs))
        # queries answered by the tensor-core filter, and how many of them the certificate sent to the exact re-run
        line["fast_path"] = {"queries": int(fs[0]), "exact_reruns": int(fs[1]), "rerun_rate": (fs[1] / fs[0]) if fs[0] else None}

    if not args.no_cpu and world == 1:
        from oracle import oracle as orc
        n_threads = min(16, os.cpu_count() or 1)
        st, n_s, cq, run, kind, have_ref = cpu_arm(args, rows, ids, q_host[0], n_threads)
        t = run()
        cpu_qps = cq / t * (n_s / n)
        line["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": n_threads, "kind": kind,
                                "kernels": "reference avx.cpp (oracle/_ref)" if have_ref else "scalar lane-order port",
                                "sample": f"{cq} queries x {n_s} of {n} rows, highCpu (16 shard workers), scaled by rows; "
                                          "contiguous rows, no Go map/alloc/GC => faster than the real Go path"}
        # recall@10 (edge/resultset.go:55-65) of the fp16 store vs fp32 ground truth from the oracle
        orc.use_reference_kernels(True)
        gt = orc.FlatStore(d, orc.COSINE, orc.Q_NONE)
        gt.upsert(ids, rows)
        rq = min(args.recall_queries, nq)
        gi, gs, gc = sp.BatchVertexSearch(q_host[0][:rq], k)
        rec = []
        for j in range(rq):
            wi, _ = gt.search_total_order(q_host[0][j], k, select_mode=orc.NEAREST, n_threads=os.cpu_count() or 1)
            rec.append(orc.compute_recall(wi, gi[j, :k], k))
        line["recall_at_10"] = float(np.mean(rec))
        line["recall_note"] = f"{rq} queries vs fp32 exact ground truth (oracle) over all {n} rows"
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
