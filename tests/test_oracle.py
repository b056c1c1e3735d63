"""Pins the CPU oracle (oracle/) before anything is compared against it.

The reference holds no golden vectors for FLAT search / codecs / top-K, so
the pins are: (1) the reference's own avx.cpp compiled unmodified (oracle/_ref) for the distance
arithmetic, bit-for-bit; (2) independent IEEE implementations (numpy float16) for the fp16 codec,
exhaustively; (3) hand-derived known answers for the literal f8 codec and the F1 heap semantics.
"""
import struct

import numpy as np
import pytest

DIMS = [1, 3, 7, 8, 9, 15, 16, 17, 31, 64, 100, 128, 384, 768, 1536, 3072]  # compresshelper_test.go dims + tails


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def test_lane_emulation_matches_reference_avx_bit_exact(oracle):
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    import ctypes as C
    L = oracle.lib()
    rng = _rng(0xC0177)
    for d in DIMS:
        for trial in range(20):
            a = oracle.aligned_f32(d)
            b = oracle.aligned_f32(d)
            if trial % 2 == 0:
                a[:] = rng.random(d, dtype=np.float32)
                b[:] = rng.random(d, dtype=np.float32)
            else:
                a[:] = rng.standard_normal(d).astype(np.float32) * 3
                b[:] = rng.standard_normal(d).astype(np.float32) * 3
            ap, bp = a.ctypes.data_as(oracle.f32p), b.ctypes.data_as(oracle.f32p)
            d1, n1, d2, n2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
            R.ref_cosine_similarity_dot_norm(d, ap, bp, C.byref(d1), C.byref(n1))
            L.orc_cosine_dot_norm(d, ap, bp, C.byref(d2), C.byref(n2))
            assert struct.pack("f", d1.value) == struct.pack("f", d2.value), (d, trial)
            assert struct.pack("f", n1.value) == struct.pack("f", n2.value), (d, trial)
            r1, r2 = C.c_float(), C.c_float()
            R.ref_euclidean_distance_squared(d, ap, bp, C.byref(r1))
            L.orc_l2sq(d, ap, bp, C.byref(r2))
            assert struct.pack("f", r1.value) == struct.pack("f", r2.value), (d, trial)
            # norm2 of one operand == the na the cosine kernel computes when nb's operand is anything
            na = oracle.norm2_avx_order(a)
            ones = oracle.aligned_f32(d)
            ones[:] = 1.0
            R.ref_cosine_similarity_dot_norm(d, ap, ones.ctypes.data_as(oracle.f32p), C.byref(d1), C.byref(n1))
            nb_ones = oracle.norm2_avx_order(ones)
            assert np.float32(n1.value) == np.float32(na) * np.float32(nb_ones)


def test_reference_kernel_hook_changes_nothing(oracle):
    rng = _rng(7)
    st = oracle.FlatStore(128, oracle.COSINE, oracle.Q_NONE)
    ids = np.arange(1, 2001, dtype=np.uint64) * 7919
    vecs = rng.random((2000, 128), dtype=np.float32)
    st.upsert(ids, vecs)
    q = rng.random(128, dtype=np.float32)
    oracle.use_reference_kernels(False)
    i0, s0 = st.search(q, 10)
    had = oracle.use_reference_kernels(True)
    i1, s1 = st.search(q, 10)
    oracle.use_reference_kernels(False)
    if had:
        assert np.array_equal(i0, i1) and s0.tobytes() == s1.tobytes()


def test_f16_codec_exhaustive_decode_and_roundtrip(oracle):
    codes = np.arange(65536, dtype=np.uint16)
    dec = oracle.f16_to_f32(codes)
    want = codes.view(np.float16).astype(np.float32)
    nan = np.isnan(want)
    assert np.array_equal(dec[~nan].view(np.uint32), want[~nan].view(np.uint32))
    assert np.all(np.isnan(dec[nan]))
    # Frombits(Bits(x)) == x for every non-NaN code (float16.go:118-120)
    back = oracle.f32_to_f16(dec)
    assert np.array_equal(back[~nan], codes[~nan])


def test_f16_encode_rne_against_numpy(oracle):
    rng = _rng(11)
    bits = rng.integers(0, 2**32, size=2_000_000, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    # plus the rounding boundaries: halfway points between adjacent fp16 values, subnormals, overflow
    h = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float64)
    mid = ((h[:-1] + h[1:]) / 2).astype(np.float32)
    x = np.concatenate([x, mid, -mid, np.nextafter(mid, np.float32(np.inf)), np.nextafter(mid, np.float32(-np.inf)),
                        np.array([65504.0, 65519.99, 65520.0, 1e6, 5.96e-8, 2.98e-8, 2.9802322e-8, 0.0, -0.0],
                                 dtype=np.float32)])
    finite = np.isfinite(x)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = oracle.f32_to_f16(x)
    assert np.array_equal(got[finite], want[finite])


def test_f8_codec_is_the_broken_reference_codec(oracle):
    """all 256 codes decode to 8 distinct values; encode keeps the low byte of the fp16 code."""
    codes = np.arange(256, dtype=np.uint8)
    dec = oracle.f8_to_f32(codes)
    # hand derivation of float8.go:233-266: exp is always 0; coef=(in&3)<<13; sign=(in&0x80)<<8
    want = np.zeros(256, dtype=np.uint32)
    for c in range(256):
        sign = (c & 0x80) << 8
        coef = (c & 0x03) << 13
        if coef == 0:
            want[c] = sign
            continue
        exp = 1
        while coef & 0x7f800000 == 0:
            coef <<= 1
            exp -= 1
        coef &= 0x007fffff
        want[c] = (sign | (((exp + (0x7f - 0xf)) & 0xffffffff) << 23) | coef) & 0xffffffff
    assert np.array_equal(dec.view(np.uint32), want)
    assert len(np.unique(dec.view(np.uint32))) == 8
    # encode: low 8 bits of the fp16 code for normal-range positives (float8.go:306-312, sign from bit 23)
    x = np.array([0.5, 0.75, 1.0, 0.1234, 0.999], dtype=np.float32)
    f16 = oracle.f32_to_f16(x)
    f8 = oracle.f32_to_f8(x)
    sign_bit23 = ((x.view(np.uint32) & 0x800000) >> 8).astype(np.uint32)
    assert np.array_equal(f8, ((f16.astype(np.uint32) | sign_bit23) & 0xff).astype(np.uint8))


def test_normalize_is_sequential_f32(oracle):
    rng = _rng(3)
    for d in (1, 5, 128, 768):
        v = rng.standard_normal(d).astype(np.float32)
        acc = np.float32(0)
        for x in v:
            acc = np.float32(acc + np.float32(x * x))
        n = np.float32(np.sqrt(np.float64(acc)))
        want = (v / n).astype(np.float32)
        got = oracle.normalize(v)
        assert got.tobytes() == want.tobytes()
    assert np.all(oracle.normalize(np.zeros(16, np.float32)) == 0)


def test_shard_vertex_is_fnv1a_le(oracle):
    def fnv(x, c):
        h = 14695981039346656037
        for i in range(8):
            h ^= (x >> (8 * i)) & 0xff
            h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h % c
    for x in (0, 1, 2, 12345678901234567, 2**64 - 1):
        assert oracle.shard_vertex(x, 16) == fnv(x, 16)


def test_edge_flat_keeps_k_largest_distances_ascending(oracle):
    """SURVEY.md F1 — edge/priority_queue.go:39-69 over a min-heap keeps the K LARGEST scores."""
    rng = _rng(5)
    n, d, k = 3000, 128, 10
    ids = rng.permutation(np.arange(10_000, 10_000 + n)).astype(np.uint64)
    vecs = rng.random((n, d), dtype=np.float32)
    q = rng.random(d, dtype=np.float32)
    for quant in (oracle.Q_NONE, oracle.Q_F16, oracle.Q_BF16):
        for metric in (oracle.COSINE, oracle.EUCLIDEAN):
            st = oracle.FlatStore(d, metric, quant)
            st.upsert(ids, vecs)
            # brute force, same arithmetic
            qn = oracle.normalize(q) if metric == oracle.COSINE else q
            if quant != oracle.Q_NONE:
                qn = oracle.f16_to_f32(oracle.f32_to_f16(qn))
            sc = np.empty(n, np.float32)
            for i in range(n):
                row = st.get_row(int(ids[i]))
                row = oracle.f16_to_f32(row) if quant != oracle.Q_NONE else row
                sc[i] = (oracle.cosine_distance if metric == oracle.COSINE else oracle.euclidean_distance)(qn, row)
            order = np.lexsort((ids, sc))
            got_ids, got_sc = st.search(q, k)
            assert np.all(np.diff(got_sc) >= 0)
            assert got_sc.tobytes() == sc[order[-k:]].tobytes()
            assert set(got_ids.tolist()) == set(ids[order[-k:]].tolist())
            for high in (False, True):
                i2, s2 = st.search(q, k, high_cpu=high, n_threads=4)
                assert s2.tobytes() == got_sc.tobytes() and np.array_equal(i2, got_ids)
            # NEAREST extension keeps the K smallest
            ni, ns = st.search(q, k, select_mode=oracle.NEAREST)
            assert ns.tobytes() == sc[order[:k]].tobytes() and np.array_equal(ni, ids[order[:k]])
            # total-order selector == literal heap selector on tie-free data
            for mode in (oracle.COLTT_COMPAT, oracle.NEAREST):
                a = st.search(q, k, select_mode=mode)
                b = st.search_total_order(q, k, select_mode=mode)
                assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()


def test_filterable_search_and_edge_cases(oracle):
    rng = _rng(9)
    n, d = 500, 32
    ids = np.arange(1, n + 1, dtype=np.uint64) * 3
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    st = oracle.FlatStore(d, oracle.COSINE, oracle.Q_NONE)
    assert len(st.search(vecs[0], 5)[0]) == 0  # empty store
    st.upsert(ids, vecs)
    cand = np.concatenate([ids[::7], np.array([999_999], dtype=np.uint64)])  # one id not present
    a = st.search_subset(vecs[3], cand, 5)
    b = st.search_total_order(vecs[3], 5, cand_ids=cand)
    assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()
    assert len(st.search(vecs[0], 10_000)[0]) == n  # K > N
    st.remove(ids[:100])
    assert len(st) == n - 100
    st.upsert(ids[100:101], vecs[0:1])  # overwrite keeps the id
    assert len(st) == n - 100


def test_resultset_semantics(oracle):
    import ctypes as C
    L = oracle.lib()
    rs = L.orc_resultset_create(3)
    for id_, sim in [(1, 0.1), (2, 0.9), (3, 0.5), (4, 0.7), (2, 0.95), (5, 0.05)]:
        L.orc_resultset_add(rs, id_, sim)
    ids = np.zeros(3, np.uint64)
    sims = np.zeros(3, np.float32)
    n = L.orc_resultset_to_slice(rs, ids.ctypes.data_as(oracle.u64p), sims.ctypes.data_as(oracle.f32p))
    # hand trace of edge/resultset.go:71-108: `valid` only grows on an append, id 2 is de-duplicated at
    # slot 0, and the late (5, 0.05) overwrites the shifted-out tail — literal reference behaviour.
    assert n == 3 and ids.tolist() == [2, 4, 5]
    assert np.allclose(sims, [0.9, 0.7, 0.05])
    L.orc_resultset_destroy(rs)
    assert oracle.compute_recall([1, 2, 3, 4], [4, 3, 9, 9], 4) == 0.5


def test_hnsw_build_search_commit_roundtrip(oracle):
    """Mirrors core/vectorindex/hnsw_commit_test.go:127-181 (random 1000x128, 20% deletes,
    Commit -> Load, equality) and e2e/hnsw/e2e_hnsw.go (ascending distances, nearest first)."""
    rng = _rng(13)
    n, d = 1000, 128
    ids = np.arange(1, n + 1, dtype=np.uint64)
    vecs = rng.random((n, d), dtype=np.float32)
    h = oracle.Hnsw(d, oracle.COSINE)
    h.build(ids, vecs)
    for i in rng.permutation(n)[: n // 5]:
        assert h.remove(int(ids[i])) == 0
    blob = h.commit()
    h2 = oracle.Hnsw.load(blob)
    assert len(h2) == len(h) == n - n // 5
    assert h2.commit() == blob
    h.set_ef(64)
    h2.set_ef(64)
    flat = oracle.FlatStore(d, oracle.COSINE, oracle.Q_NONE)
    alive = np.array([i for i in ids if True], dtype=np.uint64)
    hits = 0
    for t in range(20):
        q = rng.random(d, dtype=np.float32)
        a = h.search(q, 10)
        b = h2.search(q, 10)
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()
        assert np.all(np.diff(a[1]) >= 0)
    # recall vs brute force on a clean (no deletes) index
    h3 = oracle.Hnsw(d, oracle.COSINE, ef=64)
    h3.build(ids, vecs)
    flat.upsert(ids, vecs)
    for t in range(20):
        q = rng.random(d, dtype=np.float32)
        hi, _ = h3.search(q, 10)
        fi, _ = flat.search(q, 10, select_mode=oracle.NEAREST)
        hits += len(set(hi.tolist()) & set(fi.tolist()))
    assert hits / 200 > 0.9


def test_fp16_products_are_exact_in_fp32(oracle):
    """Premise of the one-FFMA dot step on the fp16 stores (csrc/exact_math.cuh dot_step): the product of two
    decoded fp16 values is exactly representable in fp32, so fma(q, r, acc) == (q*r rounded) + acc,
    i.e. the reference's unfused `dot += v1*v2` (avx.cpp:60).  Checked exhaustively over exponent extremes and on
    2M random pairs: the fp64 product (always exact for 11-bit x 11-bit significands) survives a round trip to fp32."""
    all16 = np.arange(65536, dtype=np.uint16).view(np.float16)
    finite = all16[np.isfinite(all16)].astype(np.float64)
    r = np.random.Generator(np.random.Philox(7))
    a = r.choice(finite, 2_000_000)
    b = r.choice(finite, 2_000_000)
    # extremes: smallest subnormals, largest normals, against everything
    ext = np.array([2.0 ** -24, 3 * 2.0 ** -24, 1023 * 2.0 ** -24, 2.0 ** -14, 65504.0, 2047 * 2.0 ** -10], np.float64)
    a = np.concatenate([a, np.repeat(ext, len(finite))])
    b = np.concatenate([b, np.tile(finite, len(ext))])
    prod = a * b
    with np.errstate(over="raise", under="raise"):
        p32 = prod.astype(np.float32)
    assert np.array_equal(p32.astype(np.float64), prod)
    assert np.all((prod == 0) | (np.abs(prod) >= 2.0 ** -126))       # never an fp32 subnormal
    # ... and the f8-compat decoder does NOT have the property (stray mantissa bit, fp32 subnormals for codes >= 0x80):
    # that store keeps the two-rounding path
    dec = np.asarray(oracle.f8_to_f32(np.arange(256, dtype=np.uint8)), np.float32)
    assert not np.array_equal(dec.astype(np.float16).astype(np.float32), dec)


def test_cflat_multi_search_is_the_sequential_f32_weighted_sum(oracle):
    """experimental MultiVertexSearch (multi_vector_vertex.go:85-137): score = sum over included fields, in request
    order, of scoreHelper(Distance) * (float32(ratio)/100) in float32; topK largest, descending — against a numpy
    float32 emulation of the Go expression, both metrics, an excluded field, and k > n."""
    r = np.random.default_rng(1)
    n, d = 400, 24
    ids = np.arange(1, n + 1, dtype=np.uint64) * 7
    F = {"title": r.standard_normal((n, d)).astype(np.float32), "body": r.standard_normal((n, d)).astype(np.float32),
         "tags": r.standard_normal((n, d)).astype(np.float32)}
    q = {f: r.standard_normal(d).astype(np.float32) for f in F}
    f32 = np.float32
    for metric in (0, 1):
        for inc, k in (([("title", 30), ("body", 70)], 5), ([("tags", 100)], 7), ([("body", 20), ("tags", 45), ("title", 35)], n + 3)):
            gi, gs = oracle.multi_search(d, metric, ids, F, [(f, q[f], ra) for f, ra in inc], k)
            sc = []
            for row in range(n):
                s = f32(0)
                for f, ra in inc:
                    if metric == 0:
                        dist = oracle.cosine_distance(oracle.normalize(F[f][row]), oracle.normalize(q[f]))
                        h = f32(f32(f32(f32(2) - f32(dist)) / f32(2)) * f32(100))
                    else:
                        dist = oracle.euclidean_distance(F[f][row], q[f])
                        h = f32(max(0.0, float(f32(f32(100) - f32(dist)))))
                    s = f32(s + f32(h * f32(f32(ra) / f32(100))))
                sc.append(s)
            sc = np.array(sc, np.float32)
            o = np.lexsort((-ids.astype(np.int64), -sc))[:k]
            assert np.array_equal(gi, ids[o]) and gs.tobytes() == sc[o].tobytes(), (metric, inc)
            assert np.all(np.diff(gs) <= 0)


def test_lane_emulation_reproduces_the_committed_reference_outputs(oracle):
    """tests/golden/avx_golden.json holds what the reference's own compiled AVX kernels returned for seeded inputs
    (generated by tests/golden/make_avx_golden.py where /root/reference exists); the oracle's scalar lane-order
    restatement must reproduce those bits with the reference hook OFF — this pins it where oracle/_ref is absent too."""
    import ctypes as C
    import json
    import os
    from tests.golden.make_avx_golden import bits, inputs
    oracle.use_reference_kernels(False)
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "avx_golden.json")))
    L = oracle.lib()
    assert len(gold["cases"]) >= 40
    for c in gold["cases"]:
        a, b = inputs(c["dim"], c["trial"])
        ap, bp = a.ctypes.data_as(oracle.f32p), b.ctypes.data_as(oracle.f32p)
        dot, n2, l2 = C.c_float(), C.c_float(), C.c_float()
        L.orc_cosine_dot_norm(c["dim"], ap, bp, C.byref(dot), C.byref(n2))
        L.orc_l2sq(c["dim"], ap, bp, C.byref(l2))
        assert (bits(dot.value), bits(n2.value), bits(l2.value)) == (c["dot"], c["norm2"], c["l2sq"]), c


def test_e4m3_codec_against_torch_float8(oracle):
    """The builder-defined F8_E4M3 store has no reference arithmetic (parity unpinned); its codec is pinned here against
    an independent implementation, torch.float8_e4m3fn: all 256 decodes, and RNE encodes of random values, every
    representable value, every midpoint between neighbours and their fp32 neighbours (|x| <= 448; above that we
    saturate to 448 where torch's cast yields NaN)."""
    import torch
    codes = np.arange(256, dtype=np.uint8)
    t = torch.from_numpy(codes).view(torch.float8_e4m3fn).to(torch.float32).numpy()
    d = oracle.e4m3_decode(codes)
    nan = np.isnan(t)
    assert np.array_equal(nan, np.isnan(d)) and np.array_equal(t[~nan].view(np.uint32), d[~nan].view(np.uint32))
    r = np.random.Generator(np.random.Philox(7))
    x = r.standard_normal(100_000).astype(np.float32) * r.choice(np.array([1e-3, 1e-2, .1, 1, 10, 100], np.float32), 100_000)
    vals = np.sort(np.unique(d[~nan]))
    mids = ((vals[:-1].astype(np.float64) + vals[1:]) / 2).astype(np.float32)
    x = np.concatenate([x, vals, mids, np.nextafter(mids, np.float32(1e9)), np.nextafter(mids, np.float32(-1e9))])
    x = x[np.abs(x) <= 448]
    want = torch.from_numpy(x).to(torch.float8_e4m3fn).view(torch.uint8).numpy()
    assert np.array_equal(oracle.e4m3_encode(x), want)
    assert oracle.e4m3_encode(np.float32([1e9, -1e9, 460.0]))[:3].tolist() == [0x7e, 0xfe, 0x7e]   # saturating


def test_e4m3_lowering_scale_and_store(oracle):
    """Lower(v): s = 2^clamp(floor(log2 max|v|) - 7, -40, 40) puts max|v|/s in [128, 256); zero / non-finite vectors use 1."""
    r = np.random.Generator(np.random.Philox(9))
    for mag in (1e-30, 1e-3, 1.0, 77.0, 1e20):
        v = (r.standard_normal(300) * mag).astype(np.float32)
        c, s = oracle.f32_to_e4m3(v)
        m = np.abs(v).max()
        lo, hi = (128, 256) if 2.0 ** -40 * 128 <= m < 2.0 ** 40 * 256 else (0, np.inf)
        assert lo <= m / s < hi and np.log2(s) == np.round(np.log2(s))
        back = oracle.e4m3_to_f32(c, s)
        if lo:
            assert np.abs(back - v).max() <= m / 16 + 1e-30      # 3 mantissa bits at the top of the range
    assert oracle.f32_to_e4m3(np.zeros(8, np.float32))[1] == 1.0
    assert oracle.f32_to_e4m3(np.float32([1, np.inf, 2]))[1] == 1.0
    # the store: scores equal the reference arithmetic over the dequantized operands
    n, d = 300, 40
    vecs = r.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint64)
    st = oracle.FlatStore(d, oracle.COSINE, oracle.Q_F8_E4M3)
    st.upsert(ids, vecs)
    q = r.standard_normal(d).astype(np.float32)
    wi, ws = st.search_total_order(q, 5, select_mode=oracle.NEAREST)
    cq, sq = oracle.f32_to_e4m3(oracle.normalize(q))
    qd = oracle.e4m3_to_f32(cq, sq)
    sc = []
    for v in vecs:
        cr, sr = oracle.f32_to_e4m3(oracle.normalize(v))
        sc.append(oracle.cosine_distance(qd, oracle.e4m3_to_f32(cr, sr)))
    order = np.lexsort((ids, np.float32(sc)))[:5]
    assert np.array_equal(wi, ids[order]) and np.array_equal(ws, np.float32(sc)[order])
