"""GPU bulk construction of a core/vectorindex HNSW graph (csrc/hnsw_build.cu) and Hnsw.Commit.

The bulk build is sequential Insert with the construction search made exact (csrc/hnsw_build.cu), so:
  * where the reference's own construction search is exhaustive (n below efConstruction) the graph must equal the
    oracle's sequential Insert edge for edge, distance bits included (test_bulk_build_equals_sequential_insert);
and at sizes where Go's approximate search may miss neighbours, what the reference's format and search define:
  * the committed blob is loadable by the restatement of Hnsw.Load (oracle) and by the device loader, and the
    device search of the built graph == the oracle's search of the loaded blob, bit for bit, with the same number
    of distance evaluations (the Commit/Load round trip of hnsw_commit_test.go:127-181, across implementations);
  * every edge's stored distance is the reference's Distance() of the two stored vectors, bit for bit;
  * degree caps (mMax / mMax0), no self edges, every edge's endpoints live on that level;
  * levels / entry point follow sequential insertion (first vertex at level 0, first vertex to reach the top);
  * recall@10 against exact search.
"""
import struct

import numpy as np
import pytest

from tests.util import QUERY_SEED, assert_same_hits, normal, sparse_ids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import coltt_b200
    from coltt_b200 import _lib
    assert _lib.lib().coltt_b200_device_count() >= 1, "needs a B200"
    return coltt_b200


def parse_commit(blob):
    """Hnsw.Load(header=true) byte format (hnsw_commit.go:164-278), parsed independently in Python."""
    o = 0

    def rd(fmt):
        nonlocal o
        v = struct.unpack_from(">" + fmt, blob, o)
        o += struct.calcsize(">" + fmt)
        return v if len(v) > 1 else v[0]
    algo, lm, ef, efc, m, mmax, mmax0, dim, di = rd("IfiiiiiIB")
    cfg = dict(algo=algo, level_mult=lm, ef=ef, efc=efc, m=m, mmax=mmax, mmax0=mmax0, dim=dim, dist=di)
    if o == len(blob):
        return cfg, None, {}, {}
    ep = rd("Q")
    verts, order = {}, []
    for _ in range(16):
        for _ in range(rd("I")):
            vid, lvl = rd("Qi")
            vec = np.frombuffer(blob, dtype=">f4", count=dim, offset=o).astype(np.float32)
            o += dim * 4
            assert rd("H") == 0
            verts[vid] = (lvl, vec)
            order.append(vid)
    edges = {}
    for vid in order:
        assert rd("Q") == vid
        per = {}
        for l in range(verts[vid][0], -1, -1):
            lst = []
            for _ in range(rd("I")):
                nid, d = rd("Qf")
                lst.append((nid, d))
            per[l] = lst
        edges[vid] = per
    assert o == len(blob)
    return cfg, ep, verts, edges


@pytest.mark.parametrize("metric,n,d,m", [(0, 3000, 64, 16), (1, 2500, 100, 8), (0, 6000, 128, 16), (1, 5000, 96, 12)])
def test_bulk_build_commit_load_search(cb, oracle, metric, n, d, m):
    ids, vecs = sparse_ids(n, n + d), normal(n, d, n + d)
    g = cb.Hnsw.Build(ids, vecs, metric=metric, m=m, seed=n)
    assert g.Len() == n
    blob = g.Commit()
    cfg, ep, verts, edges = parse_commit(blob)
    assert (cfg["m"], cfg["mmax"], cfg["mmax0"], cfg["dim"], cfg["dist"], cfg["ef"], cfg["efc"]) == (m, m, 2 * m, d, 1 if metric == 0 else 2, 20, 200)
    assert np.float32(cfg["level_mult"]) == np.float32(1.0) / np.float32(np.log(np.float64(np.float32(m))))
    assert set(verts) == set(int(i) for i in ids)
    # levels and entry point as sequential insertion leaves them
    lv = np.array([verts[int(i)][0] for i in ids])
    assert lv[0] == 0
    top = lv.max()
    assert ep == int(ids[int(np.argmax(lv == top))]) if top > 0 else ep == int(ids[0])
    assert 0.03 < (lv >= 1).mean() < 0.25                     # P(level >= 1) = 1/m
    # stored vectors are the reference's Normalize(value) / value
    for j in (0, 1, n // 2, n - 1):
        want = oracle.normalize(vecs[j]) if metric == 0 else vecs[j]
        assert verts[int(ids[j])][1].tobytes() == np.asarray(want, np.float32).tobytes()
    # edges: caps, no self loops, endpoints on the level, distances = Distance(stored, stored) bit for bit
    dist = oracle.cosine_distance if metric == 0 else oracle.euclidean_distance
    checked = 0
    for vid, per in edges.items():
        for l, lst in per.items():
            assert len(lst) <= (2 * m if l == 0 else m)
            nbrs = [x for x, _ in lst]
            assert vid not in nbrs and len(set(nbrs)) == len(nbrs)
            for nid, dd in lst[:2] if checked < 4000 else []:
                assert verts[nid][0] >= l
                want = dist(verts[vid][1], verts[nid][1])
                assert np.float32(dd).tobytes() == np.float32(want).tobytes(), (vid, nid, dd, want)
                checked += 1
    assert checked > 1000
    # level 0: vertex i chose min(m, i) predecessors; back edges only add, pruning keeps at least that many
    deg0 = np.array([len(edges[int(i)][0]) for i in ids])
    assert np.all(deg0[1:] >= np.minimum(m, np.arange(1, n))) and deg0.max() <= 2 * m

    # search parity: device search of the built graph == oracle search of the committed blob
    h = oracle.Hnsw.load(blob)
    assert len(h) == n
    g2 = cb.Hnsw.Load(blob)
    qs = normal(16, d, QUERY_SEED + d)
    for ef, k in [(0, 10), (64, 10), (128, 5)]:
        h.set_ef(ef if ef else 20)
        h.stats(reset=True)
        want = [h.search(q, k) for q in qs]
        evals, exps = h.stats(reset=True)
        for gg in (g, g2):
            gi, gs, gc = gg.BatchSearch(qs, k, ef)
            for j in range(len(qs)):
                assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], want[j][0], want[j][1], f"ef={ef} k={k} q{j}")
            st = gg.last_stats()
            assert st["dist_evals"] == evals and st["expansions"] == exps, (st, evals, exps)
    # Commit -> Load -> Commit is the identity
    assert g2.Commit() == blob

    # recall@10 of the bulk-built graph against exact search (oracle FLAT fp32, NEAREST)
    gt = oracle.FlatStore(d, metric, oracle.Q_NONE)
    gt.upsert(ids, vecs)
    gi, gs, gc = g.BatchSearch(qs, 10, 128)
    rec = []
    for j in range(len(qs)):
        wi, _ = gt.search_total_order(qs[j], 10, select_mode=oracle.NEAREST)
        rec.append(oracle.compute_recall(wi, gi[j, :10], 10))
    assert np.mean(rec) >= 0.9, rec
    g.close()
    g2.close()


@pytest.mark.parametrize("metric,n,d,m,seed", [(0, 150, 16, 8, 1), (1, 120, 24, 8, 2), (0, 180, 33, 16, 3), (1, 150, 24, 10, 4), (0, 190, 48, 12, 8)])
def test_bulk_build_equals_sequential_insert(cb, oracle, metric, n, d, m, seed):
    """n < efConstruction (200): the reference's searchLevel(efConstruction) visits every reachable vertex, so
    sequential Insert connects each vertex to its true m nearest predecessors — the graph the bulk build defines.
    Same levels in, same edges (and edge distance bits) out.  (With very small m the reference's pruning leaves
    vertices unreachable from the entry point — edges are directed after pruneNeighbors — and its search then
    misses true neighbours even at this size; the configurations here are ones where it does not.)"""
    ids, vecs = sparse_ids(n, seed), normal(n, d, seed)
    h = oracle.Hnsw(d, metric, m=m)
    us = np.maximum(np.random.Generator(np.random.Philox(seed)).random(n, dtype=np.float32), np.float32(1e-30))
    lv = np.array([h.level_from_uniform(float(u)) for u in us], np.int32)
    for i in range(n):
        h.insert(int(ids[i]), vecs[i], int(lv[i]))
    cfg_o, ep_o, verts_o, edges_o = parse_commit(h.commit())
    g = cb.Hnsw.Build(ids, vecs, metric=metric, m=m, levels=lv)
    cfg_g, ep_g, verts_g, edges_g = parse_commit(g.Commit())
    assert cfg_g == cfg_o and ep_g == ep_o
    assert {k: v[0] for k, v in verts_g.items()} == {k: v[0] for k, v in verts_o.items()}
    for vid in verts_o:
        assert verts_g[vid][1].tobytes() == verts_o[vid][1].tobytes()
        for l in edges_o[vid]:
            want = sorted((nid, np.float32(dd).tobytes()) for nid, dd in edges_o[vid][l])
            got = sorted((nid, np.float32(dd).tobytes()) for nid, dd in edges_g[vid][l])
            assert got == want, (vid, l, edges_g[vid][l], edges_o[vid][l])
    g.close()


def test_bulk_build_caller_levels_and_small_cases(cb, oracle):
    d = 32
    # caller-supplied vertexLevel per Insert (hnsw.go:104): vertex 0 is forced to level 0, entry = first at the top
    n = 400
    ids, vecs = sparse_ids(n, 3), normal(n, d, 3)
    lv = np.zeros(n, np.int32)
    lv[0] = 5          # ignored: the first vertex is created at level 0 (hnsw.go:110)
    lv[7] = 2
    lv[9] = 3
    lv[50] = 3         # same level as the entry point: does not replace it (strictly greater, hnsw.go:160)
    lv[100:140] = 1
    g = cb.Hnsw.Build(ids, vecs, metric=0, levels=lv, m=8)
    cfg, ep, verts, edges = parse_commit(g.Commit())
    assert ep == int(ids[9])
    assert verts[int(ids[0])][0] == 0 and verts[int(ids[7])][0] == 2 and verts[int(ids[50])][0] == 3
    assert g.build_stats()["max_level"] == 3
    # level 3 has two members: the later one chose the earlier one, which got the back edge
    assert [x for x, _ in edges[int(ids[9])][3]] == [int(ids[50])]
    assert [x for x, _ in edges[int(ids[50])][3]] == [int(ids[9])]
    h = oracle.Hnsw.load(g.Commit())
    qs = normal(8, d, 77)
    h.set_ef(32)
    gi, gs, gc = g.BatchSearch(qs, 5, 32)
    for j in range(len(qs)):
        wi, ws = h.search(qs[j], 5)
        assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"q{j}")
    g.close()
    # empty, one and two vertices
    g0 = cb.Hnsw.Build(np.zeros(0, np.uint64), np.zeros((0, d), np.float32))
    assert g0.Len() == 0 and len(g0.Commit()) == 33
    assert g0.BatchSearch(qs[:1], 3)[2][0] == 0
    g1 = cb.Hnsw.Build(ids[:1], vecs[:1])
    gi, gs, gc = g1.BatchSearch(qs[:2], 3)
    assert gc.tolist() == [1, 1] and gi[0, 0] == ids[0]
    g2 = cb.Hnsw.Build(ids[:2], vecs[:2], metric=1)
    gi, gs, gc = g2.BatchSearch(vecs[:2], 3)
    assert gc.tolist() == [2, 2] and gi[0, 0] == ids[0] and gi[1, 0] == ids[1] and gs[0, 0] == 0.0
    # duplicate ids are refused (a second Insert would replace the first in the Go map)
    from coltt_b200._lib import ColttError
    with pytest.raises(ColttError):
        cb.Hnsw.Build(np.array([5, 5], np.uint64), vecs[:2])
    for x in (g0, g1, g2):
        x.close()


def test_large_graph_search_matches_oracle(cb, oracle):
    """A graph large enough for long walks (expansions whose neighbours are all visited, distance ties between
    distinct vertices, several CTAs per SM): bulk-built on the GPU, committed, and searched by the oracle's literal
    restatement of hnsw.go from the same blob — ids, score bits and the evaluation / expansion counts must agree."""
    n, d, lat = 150_000, 48, 6
    g0 = np.random.Generator(np.random.Philox(77))
    proj = g0.standard_normal((lat, d), dtype=np.float32)
    vecs = (g0.standard_normal((n, lat), dtype=np.float32) @ proj + np.float32(0.05) * g0.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    vecs[100_000:104_000] = vecs[:4000]            # exact duplicates: equal priorities that meet pops and evictions
    ids = sparse_ids(n, 5)
    g = cb.Hnsw.Build(ids, vecs, metric=0, m=16, seed=3)
    blob = g.Commit()
    h = oracle.Hnsw.load(blob)
    qs = (g0.standard_normal((320, lat), dtype=np.float32) @ proj).astype(np.float32)
    qs[:40] = vecs[:40]                            # queries next to duplicated vertices
    for ef, k in [(128, 10), (200, 10), (48, 5)]:
        h.set_ef(ef)
        h.stats(reset=True)
        want = [h.search(q, k) for q in qs]
        evals, exps = h.stats(reset=True)
        gi, gs, gc = g.BatchSearch(qs, k, ef)
        for j in range(len(qs)):
            assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], want[j][0], want[j][1], f"large ef={ef} q{j}")
        st = g.last_stats()
        assert st["dist_evals"] == evals and st["expansions"] == exps, (ef, st, evals, exps)
    g.close()
