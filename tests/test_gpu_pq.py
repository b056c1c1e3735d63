"""K6 — product quantization for the HNSW walk (BASELINE config 5).  PARITY UNPINNED: the reference holds no PQ arithmetic
(SURVEY F5), so this is judged on recall against exact search, with two hard checks: re-ranked scores are the exact
distances (bit-equal to the FLAT exact kernel for the same id), and PQ recall stays close to the fp32 walk's."""
import numpy as np
import pytest

from tests.util import QUERY_SEED, rng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import coltt_b200
    from coltt_b200 import _lib
    assert _lib.lib().coltt_b200_device_count() >= 1, "needs a B200"
    return coltt_b200


def _latent(n, d, seed, lat=16):
    A = rng(0xA).standard_normal((lat, d)).astype(np.float32)
    g = rng(seed)
    return (g.standard_normal((n, lat)).astype(np.float32) @ A + np.float32(0.1) * g.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


@pytest.mark.parametrize("metric", [0, 1])
def test_pq_walk_recall_and_exact_rerank_scores(cb, oracle, metric):
    n, d, k, ef, nq = 30_000, 128, 10, 128, 64
    rows, qs = _latent(n, d, 1), _latent(nq, d, QUERY_SEED)
    ids = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(7919)
    h = cb.Hnsw.Build(ids, rows, metric=metric, m=16, ef=ef)
    with pytest.raises(cb.ColttError):
        h.BatchSearchPQ(qs, k, ef)                      # no quantizer yet
    with pytest.raises(cb.ColttError):
        h.TrainPQ(256, 7, 4096)                         # 7 does not divide 128
    h.TrainPQ(num_centroids=256, num_sub_vectors=32, trigger_threshold=8192)
    sp = cb.VectorSpace("gt", cb.Metadata(d, metric, cb.Quantization_None), select_mode=cb.SELECT_NEAREST)
    sp.ChangedVertices(ids, rows)
    wi, ws, _ = sp.BatchVertexSearch(qs, k, math_mode=cb.MATH_EXACT)
    gi, gs, gc = h.BatchSearch(qs, k, ef)
    pi, ps, pc = h.BatchSearchPQ(qs, k, ef, rerank=True)
    ai, as_, ac = h.BatchSearchPQ(qs, k, ef, rerank=False)
    rec = lambda got: float(np.mean([oracle.compute_recall(wi[j], got[j], k) for j in range(nq)]))
    r_walk, r_pq, r_adc = rec(gi), rec(pi), rec(ai)
    print(f"metric {metric}: recall@10 fp32 walk {r_walk:.3f}, PQ walk + exact rerank {r_pq:.3f}, PQ walk (ADC order) {r_adc:.3f}; stats {h.last_stats()}")
    assert np.all(pc == k) and np.all(ac == k)
    assert r_pq >= 0.75 and r_pq >= r_walk - 0.1, (r_walk, r_pq)
    assert r_adc >= 0.4
    # re-ranked scores are the reference's exact distances: ask the FLAT exact kernel about the same ids
    for j in range(0, nq, 9):
        ei, es, ec = sp.BatchVertexSearch(qs[j], k, candidate_ids=pi[j, :k], math_mode=cb.MATH_EXACT)
        assert np.array_equal(ei[0, :k], pi[j, :k]) and es[0, :k].tobytes() == ps[j, :k].tobytes(), f"q{j}"
        assert np.all(np.diff(ps[j, :k]) >= 0)
    h.close(); sp.close()
