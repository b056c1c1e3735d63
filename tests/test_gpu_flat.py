"""GPU parity tests of the FLAT path, through the C-ABI, against the CPU oracle.

Bar (BASELINE.json north_star): returned indices bit-exact, fp32 scores bit-exact in
COLTT_MATH_EXACT (tolerance 0 — tighter than the 1e-5 relative the spec allows), for
none/f16/"bf16"/f8 stores, cosine and euclidean, both select modes.
"""
import numpy as np
import pytest

from tests.util import BASE_SEED, QUERY_SEED, assert_same_hits, normal, rng, score_bits, sparse_ids, uniform

pytestmark = pytest.mark.gpu

QUANTS = [0, 1, 3, 2]   # None, F16, BF16 (== F16), F8 (literal broken codec, F3)
METRICS = [0, 1]


@pytest.fixture(scope="module")
def cb():
    import coltt_b200
    from coltt_b200 import _lib
    assert _lib.lib().coltt_b200_device_count() >= 1, "needs a B200"
    return coltt_b200


def _pair(cb, oracle, d, metric, quant, ids, vecs):
    sp = cb.VectorSpace("t", cb.Metadata(d, metric, quant))
    sp.ChangedVertices(ids, vecs)
    st = oracle.FlatStore(d, metric, quant)
    st.upsert(ids, vecs)
    return sp, st


def test_ingest_normalize_and_lower_are_bit_exact(cb, oracle):
    """ChangedVertex = Normalize (edge/vectorstore.go:173-189) + Lower: stored rows must equal the
    oracle's bits, incl. fp16 subnormals / overflow and the literal f8 code."""
    d, n = 96, 600
    v = normal(n, d) * np.float32(3)
    v[0, :] = 0                                              # zero vector stays zero
    v[1, :8] = [65504.0, 65520.0, 1e6, 5.96e-8, 2.98e-8, -2.9802322e-8, 1e-10, -70000.0]
    v[2] = np.float32(1e-6) * v[2]                           # tiny values -> fp16 subnormals
    ids = sparse_ids(n)
    for quant in QUANTS:
        for metric in METRICS:
            sp, st = _pair(cb, oracle, d, metric, quant, ids, v)
            for i in list(range(8)) + [n // 2, n - 1]:
                got, want = sp.stored_row(int(ids[i])), st.get_row(int(ids[i]))
                assert got.tobytes() == want.tobytes(), (quant, metric, i)
            sp.close()


@pytest.mark.parametrize("quant", QUANTS)
@pytest.mark.parametrize("metric", METRICS)
def test_config1_shape_parity(cb, oracle, quant, metric):
    """BASELINE config 1: edge FLAT cosine dim=128 N=10k top-10 single query (all stores, both metrics)."""
    n, d, k = 10_000, 128, 10
    ids, vecs = sparse_ids(n), uniform(n, d)
    qs = uniform(4, d, QUERY_SEED)
    sp, st = _pair(cb, oracle, d, metric, quant, ids, vecs)
    for q in qs:
        for mode in (cb.SELECT_COMPAT, cb.SELECT_NEAREST):
            hits = sp.VertexSearch(q, k, select_mode=mode)
            gi, gs = [h.Id for h in hits], np.array([h.Score for h in hits], np.float32)
            wi, ws = st.search_total_order(q, k, select_mode=mode)
            assert_same_hits(gi, gs, wi, ws, f"q={quant} m={metric} mode={mode}")
            if quant != 2:  # tie-free data: the literal Go-heap restatement agrees too
                li, ls = st.search(q, k, select_mode=mode)
                assert_same_hits(gi, gs, li, ls, "literal heap")
                hi, hs = st.search(q, k, high_cpu=True, select_mode=mode, n_threads=4)
                assert_same_hits(gi, gs, hi, hs, "literal heap, highCpu")
    sp.close()


@pytest.mark.parametrize("d", [1, 3, 7, 8, 9, 17, 100, 130, 257])
def test_ragged_dims_and_tiny_stores(cb, oracle, d):
    """AVX tail (len % 8) handling (avx.cpp:27-31,68-72), rows not a multiple of 16 B, N < one row group."""
    for quant in QUANTS:
        for metric in METRICS:
            for n in (1, 5, 16, 17, 333):
                ids, vecs = sparse_ids(n, d * 1000 + n), normal(n, d, d * 7 + n)
                sp, st = _pair(cb, oracle, d, metric, quant, ids, vecs)
                q = normal(1, d, QUERY_SEED + d)[0]
                for mode in (0, 1):
                    hits = sp.VertexSearch(q, 7, select_mode=mode)
                    wi, ws = st.search_total_order(q, 7, select_mode=mode)
                    assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, f"d={d} n={n} q={quant} m={metric}")
                sp.close()


def test_batched_queries_match_single_queries(cb, oracle):
    """The batch surface (QT = 4 and 8 query tiles, several passes) returns what nq single calls return."""
    n, d, k = 5000, 768, 10
    ids, vecs = sparse_ids(n), normal(n, d)
    for quant, metric in [(0, 0), (3, 0), (1, 1), (2, 0)]:
        sp, st = _pair(cb, oracle, d, metric, quant, ids, vecs)
        for nq in (3, 8, 19):
            qs = normal(nq, d, QUERY_SEED + nq)
            for mode in (0, 1):
                gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=mode)
                for j in range(nq):
                    wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
                    assert gc[j] == len(wi)
                    assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"nq={nq} j={j} quant={quant}")
        sp.close()


def test_topk_sizes_and_empty(cb, oracle):
    n, d = 3000, 64
    ids, vecs = sparse_ids(n), normal(n, d)
    sp, st = _pair(cb, oracle, d, 0, 0, ids, vecs)
    q = normal(1, d, QUERY_SEED)[0]
    for k in (1, 2, 31, 32, 33, 100, 1000, 1024):
        for mode in (0, 1):
            hits = sp.VertexSearch(q, k, select_mode=mode)
            wi, ws = st.search_total_order(q, k, select_mode=mode)
            assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, f"k={k}")
    sp.close()
    # K > N and the empty store (VertexSearch on an empty map returns an empty slice)
    sp2, st2 = _pair(cb, oracle, d, 0, 0, ids[:40], vecs[:40])
    hits = sp2.VertexSearch(q, 100)
    wi, ws = st2.search_total_order(q, 100, select_mode=0)
    assert len(hits) == 40
    assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, "k>n")
    sp2.close()
    sp3 = cb.VectorSpace("e", cb.Metadata(d))
    assert sp3.VertexSearch(q, 10) == []
    sp3.close()


def test_ties_follow_the_documented_total_order(cb, oracle):
    """Duplicate vectors and the f8 store (8 decodable values => massive score ties): ids must follow
    the (score,id) window rule; scores must equal the literal reference heap's multiset regardless."""
    n, d, k = 2000, 64, 25
    base = normal(50, d)
    vecs = base[rng(1).integers(0, 50, size=n)]     # every vector appears ~40 times
    ids = sparse_ids(n)
    q = normal(1, d, QUERY_SEED)[0]
    for quant in (0, 1, 2):
        sp, st = _pair(cb, oracle, d, 0, quant, ids, vecs)
        for mode in (0, 1):
            hits = sp.VertexSearch(q, k, select_mode=mode)
            wi, ws = st.search_total_order(q, k, select_mode=mode)
            assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, f"ties quant={quant}")
            _, ls = st.search(q, k, select_mode=mode)       # literal Go heap: same scores, tie ids may differ
            assert score_bits([h.Score for h in hits]) == score_bits(ls)
        sp.close()


def test_filterable_search_is_the_same_scan_behind_a_gather(cb, oracle):
    n, d, k = 6000, 128, 10
    ids, vecs = sparse_ids(n), uniform(n, d)
    r = rng(77)
    for quant, metric in [(0, 0), (1, 0), (0, 1), (2, 1)]:
        sp, st = _pair(cb, oracle, d, metric, quant, ids, vecs)
        for n_c in (1, 15, 16, 500, 3000):
            cand = r.choice(ids, size=n_c, replace=False)
            cand = np.concatenate([cand, cand[:3], np.array([12345], np.uint64)])  # repeats + an unknown id
            q = uniform(1, d, QUERY_SEED + n_c)[0]
            for mode in (0, 1):
                hits = sp.FilterableVertexSearch(cand, q, k, select_mode=mode)
                wi, ws = st.search_total_order(q, k, select_mode=mode, cand_ids=cand)
                assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, f"subset n_c={n_c} quant={quant}")
        assert sp.FilterableVertexSearch(np.array([999], np.uint64), uniform(1, d)[0], k) == []
        sp.close()


def test_upsert_overwrite_and_remove(cb, oracle):
    n, d, k = 4000, 96, 12
    ids, vecs = sparse_ids(n), normal(n, d)
    sp, st = _pair(cb, oracle, d, 0, 1, ids, vecs)
    new = normal(300, d, 4242)
    sp.ChangedVertices(ids[100:400], new)      # overwrite keeps ids (none_vectorstore.go:67-85)
    st.upsert(ids[100:400], new)
    drop = np.concatenate([ids[::5], np.array([7], np.uint64)])   # unknown id ignored
    sp.RemoveVertex(drop)
    st.remove(drop)
    dup_ids = np.array([ids[1], ids[1], ids[2]], np.uint64)       # last write wins inside one batch
    dup_v = normal(3, d, 99)
    sp.ChangedVertices(dup_ids, dup_v)
    st.upsert(dup_ids, dup_v)
    assert sp.LoadSize() == len(st)
    for q in normal(3, d, QUERY_SEED):
        for mode in (0, 1):
            hits = sp.VertexSearch(q, k, select_mode=mode)
            wi, ws = st.search_total_order(q, k, select_mode=mode)
            assert_same_hits([h.Id for h in hits], [h.Score for h in hits], wi, ws, "after mutations")
    with pytest.raises(ValueError, match="Dim Length UnmatchdError"):
        sp.ChangedVertex("", 1, np.zeros(d + 1, np.float32))
    sp.close()


def test_save_load_vertex_blob_roundtrip(cb, oracle):
    """SaveVertex/LoadVertex (none_vectorstore.go:308-516): the GPU store writes the reference's blob
    byte for byte (same shard/ascending-id order as the oracle writer) and reloads it."""
    n, d = 700, 48
    ids, vecs = sparse_ids(n), normal(n, d)
    for quant in QUANTS:
        sp, st = _pair(cb, oracle, d, 0, quant, ids, vecs)
        blob = sp.SaveVertex()
        assert blob == st.save_vertex(), f"blob differs for quant={quant}"
        sp2 = cb.VectorSpace("l", cb.Metadata(d, 0, quant))
        sp2.LoadVertex(blob)
        assert sp2.LoadSize() == n
        q = normal(1, d, QUERY_SEED)[0]
        for mode in (0, 1):
            a = sp.VertexSearch(q, 9, select_mode=mode)
            b = sp2.VertexSearch(q, 9, select_mode=mode)
            assert [(h.Id, h.Score) for h in a] == [(h.Id, h.Score) for h in b]
        with pytest.raises(cb.ColttError):
            sp2.LoadVertex(blob[: len(blob) // 2])
        sp.close()
        sp2.close()


def test_config2_shape_exact_path_sample(cb, oracle):
    """BASELINE config 2 shape (FLAT "bf16" cosine dim=768, top-10) at an oracle-sized N."""
    n, d, k, nq = 60_000, 768, 10, 16
    ids, vecs = sparse_ids(n), normal(n, d)
    sp, st = _pair(cb, oracle, d, 0, 3, ids, vecs)
    qs = normal(nq, d, QUERY_SEED)
    gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=1)
    for j in range(nq):
        wi, ws = st.search_total_order(qs[j], k, select_mode=1)
        assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"c2 j={j}")
    sp.close()


def test_nan_scores_rank_last(cb, oracle):
    """Zero vectors give 0/0 cosine (NaN) in the reference too; documented rule: NaN sorts after numbers."""
    d = 32
    vecs = normal(40, d)
    vecs[5] = 0
    vecs[9] = 0
    ids = np.arange(1, 41, dtype=np.uint64)
    sp, st = _pair(cb, oracle, d, 0, 0, ids, vecs)
    q = normal(1, d, QUERY_SEED)[0]
    for mode in (0, 1):
        hits = sp.VertexSearch(q, 40, select_mode=mode)
        wi, ws = st.search_total_order(q, 40, select_mode=mode)
        assert [h.Id for h in hits] == wi.tolist()
        assert np.isnan([h.Score for h in hits][-2:]).all()
    sp.close()


def test_micro_batcher_rows_equal_single_query_calls(cb, oracle):
    """f-4: 8 threads of single-query callers coalesced by coltt_b200.batcher.MicroBatcher into batched searches —
    every caller gets exactly what its own VertexSearch call returns (= the oracle's answer)."""
    import threading
    from coltt_b200.batcher import MicroBatcher
    n, d, k = 6000, 96, 7
    ids, vecs = sparse_ids(n, 21), normal(n, d, 21)
    sp, st = _pair(cb, oracle, d, 0, 1, ids, vecs)
    qs = normal(96, d, QUERY_SEED + 5)
    mb = MicroBatcher(lambda q, kk: sp.BatchVertexSearch(q, kk, select_mode=cb.SELECT_NEAREST), d, max_batch=32, max_wait_ms=5.0)
    out = [None] * len(qs)

    def worker(lo, hi):
        for i in range(lo, hi):
            out[i] = mb.VertexSearch(qs[i], k)
    ths = [threading.Thread(target=worker, args=(i * 12, (i + 1) * 12)) for i in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    mb.close()
    assert mb.served == len(qs) and mb.batches < len(qs)
    for i in range(len(qs)):
        wi, ws = st.search_total_order(qs[i], k, select_mode=1)
        assert_same_hits(out[i][0], out[i][1], wi, ws, f"batcher q{i}")
    sp.close()


def test_config4_shape_f8_dim1536_top100_batched_and_sharded(cb, oracle):
    """BASELINE config 4 shape at an oracle-sized N: edge FLAT f8 (the reference's literal codec: 8 decodable values,
    massive ties -> total order T), cosine, dim 1536, top-100, a query batch, and the rows split into two shards by
    ShardVertex (pkg/sharding/shard.go:34-41) whose per-shard top-100 lists merge into the single-store answer."""
    n, d, k = 12_000, 1536, 100
    ids, vecs = sparse_ids(n, 44), normal(n, d, 44)
    qs = normal(12, d, QUERY_SEED + 44)
    sp, st = _pair(cb, oracle, d, 0, 2, ids, vecs)
    shard = np.array([oracle.shard_vertex(int(i), 16) % 2 for i in ids])
    parts = [_pair(cb, oracle, d, 0, 2, ids[shard == r], vecs[shard == r])[0] for r in (0, 1)]
    for mode in (cb.SELECT_COMPAT, cb.SELECT_NEAREST):
        gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=mode)
        per = [p.BatchVertexSearch(qs, k, select_mode=mode) for p in parts]
        for j in range(len(qs)):
            wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
            assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"c4 mode={mode} q{j}")
            # merge of the two shards' lists under T == the unsharded answer (what dist.py's K5 merge computes on the GPU)
            mi = np.concatenate([p[0][j, :p[2][j]] for p in per])
            ms = np.concatenate([p[1][j, :p[2][j]] for p in per])
            key = np.where(np.isnan(ms), np.inf, ms)
            order = np.lexsort((mi, np.isnan(ms), key))
            order = order[:k] if mode == cb.SELECT_NEAREST else order[-k:]
            assert_same_hits(mi[order], ms[order], wi, ws, f"c4 sharded mode={mode} q{j}")
    for p in parts + [sp]:
        p.close()


def test_mutations_wait_for_an_asynchronous_search(cb, oracle):
    """coltt_b200_store_search_dev with a caller stream only enqueues and returns.  An upsert / remove issued right after it
    edits rows, norms and ids in place on the store's own stream, so it must first wait for the outstanding search
    (Store::wait_for_searches): the search has to see the collection as it was when it was called."""
    import ctypes as C
    import torch
    from coltt_b200 import _lib
    L = _lib.lib()
    n, d, k, nq = 200_000, 256, 10, 64
    ids, vecs = sparse_ids(n), normal(n, d)
    sp = cb.VectorSpace("race", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None))
    sp.ChangedVertices(ids, vecs)
    st = oracle.FlatStore(d, oracle.COSINE, oracle.Q_NONE)
    st.upsert(ids, vecs)
    qs = normal(nq, d, QUERY_SEED)
    q = torch.from_numpy(qs).cuda()
    out = torch.empty((nq, k, 4), dtype=torch.int32, device="cuda")
    cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    for trial in range(3):
        _lib.check(L.coltt_b200_store_search_dev(sp._h, q.data_ptr(), nq, k, cb.SELECT_NEAREST, cb.MATH_EXACT, out.data_ptr(), cnt.data_ptr(),
                                                  stream.cuda_stream))          # 8 exact passes over 200 MB: still running when we return
        sp.ChangedVertices(ids[:60_000], normal(60_000, d, 100 + trial))        # overwrite in place
        sp.RemoveVertex(ids[100_000:160_000])                                   # move tail rows into the holes
        stream.synchronize()
        h = out.cpu().numpy().view(np.uint8).reshape(nq, k, 16)
        gi, gs = h[..., :8].copy().view(np.uint64)[..., 0], h[..., 8:12].copy().view(np.float32)[..., 0]
        for j in range(0, nq, 7):
            wi, ws = st.search_total_order(qs[j], k, select_mode=oracle.NEAREST)
            assert_same_hits(gi[j], gs[j], wi, ws, f"trial {trial} q{j}: the search saw a half-mutated store")
        sp.ChangedVertices(ids, vecs)                                            # restore for the next trial
    sp.close()


def test_page_locked_caller_buffers_are_searched_in_place(cb, oracle):
    """Host-pointer searches DMA straight out of a caller buffer that is page-locked (coltt_b200_host_alloc) and stage ordinary
    memory through the handle's pinned buffer: same answer either way, repeated calls (the second and later ones replay the
    cached graph, whose H2D node must not have captured the caller's pointer), and a buffer whose contents change between calls."""
    n, d, k, nq = 30_000, 768, 10, 40
    ids, vecs = sparse_ids(n), normal(n, d)
    sp, st = _pair(cb, oracle, d, 0, 3, ids, vecs)
    pin = cb.pinned_empty((nq, d), np.float32)
    assert pin.shape == (nq, d) and pin.dtype == np.float32 and pin.flags["C_CONTIGUOUS"]
    for mm in (cb.MATH_FAST, cb.MATH_EXACT):
        for trial in range(4):
            qs = normal(nq, d, QUERY_SEED + 50 + trial)
            pin[...] = qs                                  # the same pinned buffer, new contents
            gi, gs, gc = sp.BatchVertexSearch(pin, k, select_mode=cb.SELECT_NEAREST, math_mode=mm)
            pi, ps, pc = sp.BatchVertexSearch(qs, k, select_mode=cb.SELECT_NEAREST, math_mode=mm)      # pageable twin
            assert np.array_equal(gi, pi) and gs.tobytes() == ps.tobytes() and np.array_equal(gc, pc)
            for j in range(0, nq, 9):
                wi, ws = st.search_total_order(qs[j], k, select_mode=oracle.NEAREST)
                assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"mm={mm} trial={trial} j={j}")
    # a slice that does not start at the allocation's base is still inside the page-locked range
    gi, gs, gc = sp.BatchVertexSearch(pin[8:24], k, select_mode=cb.SELECT_NEAREST)
    for j in (0, 15):
        wi, ws = st.search_total_order(np.array(pin[8 + j]), k, select_mode=oracle.NEAREST)
        assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"slice j={j}")
    sp.close()
