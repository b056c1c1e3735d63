"""F8_E4M3 store (builder-defined real fp8: E4M3 codes + one power-of-two scale per row, include/coltt_b200.h) and the
top-100 / dim-1536 / batch-1024 shape of BASELINE config 4.

The reference has no arithmetic for this store (its f8 codec is broken), so reference parity is UNPINNED;
what is pinned: the codec against torch.float8_e4m3fn (tests/test_oracle.py), the GPU against the oracle's restatement
of the same definition bit for bit (ids + fp32 score bits), and COLTT_MATH_FAST (tcgen05 kind::f8f6f4 filter + exact
re-rank + certificate) against COLTT_MATH_EXACT bit for bit."""
import ctypes as C

import numpy as np
import pytest

from tests.util import QUERY_SEED, assert_same_hits, normal, sparse_ids, uniform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import coltt_b200
    from coltt_b200 import _lib
    assert _lib.lib().coltt_b200_device_count() >= 1, "needs a B200"
    return coltt_b200


def _fast_scores(sp, qs, k, mode):
    from coltt_b200 import _lib
    L = _lib.lib()
    f = L.coltt_b200_debug_fast_scores
    f.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64),
                  C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    nq, n = qs.shape[0], sp.LoadSize()
    acc = np.zeros((nq, n), np.float32)
    ids = np.zeros((nq, k), np.uint64)
    sc = np.zeros((nq, k), np.float32)
    cnt = np.zeros(nq, np.int32)
    q = np.ascontiguousarray(qs, np.float32)
    _lib.check(f(sp._h, q.ctypes.data_as(C.POINTER(C.c_float)), nq, k, mode, acc.ctypes.data_as(C.POINTER(C.c_float)),
                 ids.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_int32))))
    return acc, ids, sc, cnt


@pytest.mark.parametrize("d,metric", [(128, 0), (100, 0), (1536, 0), (37, 1), (768, 1)])
def test_e4m3_store_matches_oracle_bit_for_bit(cb, oracle, d, metric):
    """Ingest (Normalize -> scale -> E4M3) and the exact search path: stored codes, ids and score bits equal the oracle's."""
    n, k = 6000, 10
    ids = sparse_ids(n)
    vecs = normal(n, d) * np.float32(3.0)
    vecs[5] = 0.0                                   # zero row: scale 1, NaN cosine distance
    vecs[6, : d // 2] *= np.float32(1e-6)           # wide dynamic range inside one row (subnormal E4M3 codes)
    sp = cb.VectorSpace("e", cb.Metadata(d, metric, cb.Quantization_F8_E4M3))
    sp.ChangedVertices(ids, vecs)
    st = oracle.FlatStore(d, metric, oracle.Q_F8_E4M3)
    st.upsert(ids, vecs)
    for j in (0, 5, 6, 17, n - 1):
        assert np.array_equal(sp.stored_row(int(ids[j])), st.get_row(int(ids[j]))), f"stored codes of row {j}"
    qs = normal(9, d, QUERY_SEED)
    for mode in (cb.SELECT_COMPAT, cb.SELECT_NEAREST):
        gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=cb.MATH_EXACT)
        for j in range(len(qs)):
            wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
            assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"d={d} metric={metric} mode={mode} q{j}")
    # upsert-overwrite and remove keep the scale array in step with the rows
    sp.ChangedVertices(ids[:50], vecs[100:150] * np.float32(40.0))
    st.upsert(ids[:50], vecs[100:150] * np.float32(40.0))
    sp.RemoveVertex(ids[1000:1400])
    st.remove(ids[1000:1400])
    gi, gs, gc = sp.BatchVertexSearch(qs[:3], k, select_mode=cb.SELECT_NEAREST, math_mode=cb.MATH_EXACT)
    for j in range(3):
        wi, ws = st.search_total_order(qs[j], k, select_mode=oracle.NEAREST)
        assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"after mutation q{j}")
    sp.close()


@pytest.mark.parametrize("d,n,nq", [(1536, 8192, 5), (1536, 8200, 130), (768, 4096, 256), (100, 5000, 200)])
def test_fp8_tensor_core_accumulators_match_fp64_reference(cb, oracle, d, n, nq):
    """kind::f8f6f4 itself: acc[q][row] = sum_k dec(q)[k] * dec(row)[k] over the unscaled E4M3 values.  Every product is
    exact; the tolerance is the certificate margin the library uses for this dim (relative to ||q|| ||row||)."""
    from coltt_b200 import _lib
    ids = np.arange(1, n + 1, dtype=np.uint64)
    vecs = normal(n, d)
    qs = normal(nq, d, QUERY_SEED)
    sp = cb.VectorSpace("g", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_F8_E4M3))
    sp.ChangedVertices(ids, vecs)
    acc, gi, gs, gc = _fast_scores(sp, qs, 10, cb.SELECT_NEAREST)
    low = [oracle.f32_to_e4m3(oracle.normalize(v)) for v in vecs]
    rows = np.stack([oracle.e4m3_decode(c) for c, _ in low]).astype(np.float64)
    lq = [oracle.f32_to_e4m3(oracle.normalize(q)) for q in qs]
    qd = np.stack([oracle.e4m3_decode(c) for c, _ in lq]).astype(np.float64)
    want = qd @ rows.T
    assert np.isfinite(acc).all(), "some accumulators were never written"
    denom = np.linalg.norm(qd, axis=1)[:, None] * np.linalg.norm(rows, axis=1)[None, :]
    rel = np.abs(acc.astype(np.float64) - want) / denom
    eps = float(_lib.lib().coltt_b200_fast_eps_rel(d))
    print(f"fp8 d={d}: max relative accumulator error {rel.max():.3e} (certificate margin {eps:.3e})")
    assert rel.max() < eps, f"max rel err {rel.max()} >= eps {eps} at {np.unravel_index(rel.argmax(), rel.shape)}"
    ei, es, ec = sp.BatchVertexSearch(qs, 10, select_mode=cb.SELECT_NEAREST, math_mode=cb.MATH_EXACT)
    for j in range(nq):
        assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], ei[j, :ec[j]], es[j, :ec[j]], f"fast vs exact q{j}")
    sp.close()


@pytest.mark.parametrize("quant,d", [(16, 1536), (16, 768), (3, 768), (3, 200)])
@pytest.mark.parametrize("k", [10, 100])
def test_fast_equals_exact_top100_and_batch1024(cb, quant, d, k):
    """FAST == EXACT (ids + score bits) for top-10 and top-100, 1 / 256 / 1024 queries, both select modes, on the E4M3
    store (kind::f8f6f4) and the fp16 store (kind::f16)."""
    n = 100_000
    ids = sparse_ids(n)
    vecs = normal(n, d)
    sp = cb.VectorSpace("f", cb.Metadata(d, cb.Distance_Cosine, quant))
    sp.ChangedVertices(ids, vecs)
    for nq in (1, 256, 1024):
        qs = normal(nq, d, QUERY_SEED + nq)
        for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
            fi, fs, fc = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=cb.MATH_FAST)
            sel = np.arange(nq) if nq <= 32 else np.linspace(0, nq - 1, 32).astype(int)   # EXACT is 1 pass per 8 queries
            ei, es, ec = sp.BatchVertexSearch(qs[sel], k, select_mode=mode, math_mode=cb.MATH_EXACT)
            for a, j in enumerate(sel):
                assert_same_hits(fi[j, :fc[j]], fs[j, :fc[j]], ei[a, :ec[a]], es[a, :ec[a]], f"quant={quant} d={d} k={k} nq={nq} mode={mode} q{j}")
    st = sp.fast_stats()
    assert st["queries"] > 0, "FAST path was never taken"
    assert st["exact_reruns"] <= st["queries"] * 0.05, f"too many uncertified queries: {st}"
    sp.close()


def test_e4m3_top100_uniform_data_and_oracle(cb, oracle):
    """Config-4 shape (E4M3, cosine, dim 1536, top-100) on reference-style uniform[0,1) data, FAST against the ORACLE."""
    n, d, k = 100_000, 1536, 100
    ids = sparse_ids(n)
    vecs = uniform(n, d)
    sp = cb.VectorSpace("u", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_F8_E4M3))
    sp.ChangedVertices(ids, vecs)
    st = oracle.FlatStore(d, oracle.COSINE, oracle.Q_F8_E4M3)
    st.upsert(ids, vecs)
    qs = uniform(64, d, QUERY_SEED)
    for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
        fi, fs, fc = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=cb.MATH_FAST)
        for j in (0, 21, 63):
            wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
            assert_same_hits(fi[j, :fc[j]], fs[j, :fc[j]], wi, ws, f"uniform mode={mode} q{j}")
    sp.close()


def test_compat_f8_config4_shape_top100_vs_oracle(cb, oracle):
    """The reference's literal f8 codec at config 4's shape: 100 K x 1536, top-100, both select modes, bit-exact vs the oracle
    (exact kernel: the compat codec has 8 decodable values and is a parity mode, not a performance mode)."""
    n, d, k = 100_000, 1536, 100
    ids = sparse_ids(n)
    vecs = uniform(n, d)
    sp = cb.VectorSpace("c", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_F8))
    sp.ChangedVertices(ids, vecs)
    st = oracle.FlatStore(d, oracle.COSINE, oracle.Q_F8)
    st.upsert(ids, vecs)
    qs = uniform(8, d, QUERY_SEED)
    for mode in (cb.SELECT_COMPAT, cb.SELECT_NEAREST):
        gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=cb.MATH_FAST)   # FAST request: served exactly
        for j in (0, 3, 7):
            wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
            assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"compat f8 mode={mode} q{j}")
    sp.close()


def test_append_device_rows_equals_host_upsert(cb):
    """coltt_b200_store_append_dev: rows already in device memory give the same store as a host upsert (ids = id_base + slot)."""
    import torch
    n, d, k = 20_000, 256, 10
    vecs = normal(n, d)
    base = 1 << 33
    a = cb.VectorSpace("a", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_F8_E4M3))
    t = torch.from_numpy(vecs).cuda()
    a.AppendDeviceRows(t.data_ptr(), n // 2, d, base)
    a.AppendDeviceRows(t[n // 2:].data_ptr(), n - n // 2, d, base)
    b = cb.VectorSpace("b", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_F8_E4M3))
    b.ChangedVertices(np.arange(n, dtype=np.uint64) + np.uint64(base), vecs)
    qs = normal(16, d, QUERY_SEED)
    for mm in (cb.MATH_EXACT, cb.MATH_FAST):
        ai, as_, ac = a.BatchVertexSearch(qs, k, select_mode=1, math_mode=mm)
        bi, bs, bc = b.BatchVertexSearch(qs, k, select_mode=1, math_mode=mm)
        assert np.array_equal(ai, bi) and as_.tobytes() == bs.tobytes() and np.array_equal(ac, bc)
    with pytest.raises(cb.ColttError):
        a.ChangedVertices(np.array([1], np.uint64), vecs[:1])
    a.close(); b.close()


@pytest.mark.parametrize("metric", [0, 1])
def test_e4m3_edge_cases(cb, oracle, metric):
    """Tiny dims, k larger than the store, empty store, rows whose magnitudes hit the scale clamp (2^+-40), infinities
    (scale 1, saturating encode) and all-zero rows: the exact kernel equals the oracle's restatement bit for bit."""
    for d in (1, 3, 17):
        n, k = 60, 100
        vecs = normal(n, d)
        vecs[3] *= np.float32(1e30)
        vecs[4] *= np.float32(1e-30)
        vecs[5] = 0.0
        vecs[6, 0] = np.float32(np.inf)
        vecs[7] = np.float32(3.0e38)
        ids = sparse_ids(n)
        sp = cb.VectorSpace("edge", cb.Metadata(d, metric, cb.Quantization_F8_E4M3))
        st = oracle.FlatStore(d, metric, oracle.Q_F8_E4M3)
        gi, gs, gc = sp.BatchVertexSearch(normal(2, d, QUERY_SEED), k)
        assert gc.tolist() == [0, 0]                                       # empty store
        sp.ChangedVertices(ids, vecs)
        st.upsert(ids, vecs)
        for j in (0, 3, 4, 5, 6, 7, n - 1):
            assert np.array_equal(sp.stored_row(int(ids[j])), st.get_row(int(ids[j]))), f"d={d} stored codes of row {j}"
        qs = np.concatenate([normal(3, d, QUERY_SEED), vecs[3:5], np.zeros((1, d), np.float32)])
        for mode in (cb.SELECT_COMPAT, cb.SELECT_NEAREST):
            for mm in (cb.MATH_EXACT, cb.MATH_FAST):                         # FAST: below 4096 rows -> served exactly
                gi, gs, gc = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=mm)
                for j in range(len(qs)):
                    wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
                    assert gc[j] == n
                    assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"d={d} metric={metric} mode={mode} math={mm} q{j}")
        sp.close()
