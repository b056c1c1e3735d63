"""GPU parity of the CFLAT multi-vector search (experimental/multi_vector_vertex.go:85-137): one fp32 store per vector
field, one exact scan per included field accumulating scoreHelper(distance) * ratio/100 into a per-row score, fused
top-K on the last field — ids and score bits must equal the oracle's sequential float32 restatement."""
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


def _collection(cb, n, d, metric, fields, seed, data=normal):
    ids = sparse_ids(n, seed)
    mats = {f: data(n, d, seed + 17 * i) for i, f in enumerate(fields)}
    mv = cb.MultiVectorVertex("c", d, metric, fields)
    mv.ChangedVertices(ids, mats)
    return mv, ids, mats


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d", [(1, 8), (37, 20), (5000, 128), (20000, 770)])
def test_multi_vertex_search_matches_oracle(cb, oracle, metric, n, d):
    fields = ["title", "body", "tags"]
    mv, ids, mats = _collection(cb, n, d, metric, fields, n + d)
    qs = {f: normal(1, d, QUERY_SEED + i)[0] for i, f in enumerate(fields)}
    cases = [([("title", 30), ("body", 70)], 10), ([("tags", 100)], 3), ([("body", 20), ("tags", 45), ("title", 35)], 100),
             ([("title", 50), ("title", 50)], 5)]
    for inc, k in cases:
        req = [cb.MultiVectorIndex(f, qs[f], True, ra) for f, ra in inc] + [cb.MultiVectorIndex("body", qs["body"], False, 99)]
        got = mv.MultiVertexSearch(k, req)
        wi, ws = oracle.multi_search(d, metric, ids, mats, [(f, qs[f], ra) for f, ra in inc], k)
        assert_same_hits([x.Id for x in got], [x.Score for x in got], wi, ws, f"n={n} d={d} m={metric} {inc} k={k}")
        assert len(got) == min(k, n) and all(got[i].Score >= got[i + 1].Score for i in range(len(got) - 1))
    mv.close()


def test_multi_updates_removes_ties_and_errors(cb, oracle):
    n, d = 3000, 64
    fields = ["a", "b"]
    mv, ids, mats = _collection(cb, n, d, 0, fields, 9, data=uniform)
    # duplicate vertices (equal scores: id descending decides), overwrite, remove
    mats["a"][100:200] = mats["a"][:100]
    mats["b"][100:200] = mats["b"][:100]
    mv.ChangedVertices(ids[100:200], {f: mats[f][100:200] for f in fields})
    for i in (5, 17, 2999):
        mv.RemoveVertex(int(ids[i]))
    keep = np.ones(n, bool)
    keep[[5, 17, 2999]] = False
    ids2, mats2 = ids[keep], {f: mats[f][keep] for f in fields}
    assert mv.LoadSize() == n - 3
    qa, qb = mats["a"][3].copy(), mats["b"][3].copy()
    got = mv.MultiVertexSearch(20, [cb.MultiVectorIndex("a", qa, True, 60), cb.MultiVectorIndex("b", qb, True, 40)])
    wi, ws = oracle.multi_search(d, 0, ids2, mats2, [("a", qa, 60), ("b", qb, 40)], 20)
    assert_same_hits([x.Id for x in got], [x.Score for x in got], wi, ws, "dups")
    with pytest.raises(ValueError, match="sum of the ratios must be 100"):
        mv.MultiVertexSearch(5, [cb.MultiVectorIndex("a", qa, True, 60), cb.MultiVectorIndex("b", qb, True, 30)])
    with pytest.raises(ValueError, match="is not defined vector fields"):
        mv.MultiVertexSearch(5, [cb.MultiVectorIndex("zzz", qa, True, 100)])
    with pytest.raises(ValueError, match="expect dimension"):
        mv.MultiVertexSearch(5, [cb.MultiVectorIndex("a", qa[:10], True, 100)])
    mv.close()


def test_multi_scan_is_one_hbm_pass_per_field(cb):
    """1 M x 128 x 3 fields: the search costs about three FLAT scans (reported, loosely bounded)."""
    import time
    n, d = 1_000_000, 128
    fields = ["a", "b", "c"]
    mv, ids, mats = _collection(cb, n, d, 0, fields, 3)
    req = [cb.MultiVectorIndex("a", mats["a"][7], True, 50), cb.MultiVectorIndex("b", mats["b"][7], True, 30), cb.MultiVectorIndex("c", mats["c"][7], True, 20)]
    got = mv.MultiVertexSearch(10, req)
    assert got[0].Id == int(ids[7])
    t0 = time.perf_counter()
    for _ in range(10):
        mv.MultiVertexSearch(10, req)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    gbps = 3 * n * (d * 4 + 12) / (ms * 1e-3) / 1e9
    print(f"\nCFLAT 1M x 128 x 3 fields: {ms:.3f} ms per search, {gbps:.0f} GB/s algorithmic")
    assert ms < 20
    mv.close()
