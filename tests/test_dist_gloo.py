"""The N>1 host path on CPU: world_size-2 gloo.  Rows are partitioned with the reference's ShardVertex,
each rank answers over its shard (here: the oracle stands in for the GPU search — this file tests the
plumbing, i.e. partition -> all-gather layout -> merge semantics), and the merged answer must equal a
single store over all rows."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_vertex_matches_oracle(oracle):
    from coltt_b200.dist import gpu_of, shard_vertex
    ids = np.concatenate([np.arange(0, 50, dtype=np.uint64), np.array([2**64 - 1, 2**63, 123456789012345], dtype=np.uint64)])
    want = np.array([oracle.shard_vertex(int(i), 16) for i in ids])
    assert np.array_equal(shard_vertex(ids, 16), want)
    assert np.array_equal(gpu_of(ids, 4), want % 4)


def _numpy_merge(gathered, gcounts, k, select_mode):
    """T-order merge (ascending score, NaN last, then id): NEAREST keeps the first k, COMPAT the last k."""
    import torch
    from coltt_b200.dist import HIT_DTYPE
    g = gathered.contiguous().numpy().view(HIT_DTYPE).reshape(gathered.shape[0], gathered.shape[1], gathered.shape[2])
    c = gcounts.contiguous().numpy()
    nq = g.shape[1]
    out = np.zeros((nq, k), dtype=HIT_DTYPE)
    cnt = np.zeros(nq, dtype=np.int32)
    for q in range(nq):
        allh = np.concatenate([g[r, q, : c[r, q]] for r in range(g.shape[0])])
        nan = np.isnan(allh["score"])
        order = np.lexsort((allh["id"], np.where(nan, np.inf, allh["score"]), nan))
        allh = allh[order]
        sel = allh[:k] if select_mode == 1 else allh[max(0, len(allh) - k):]
        out[q, : len(sel)] = sel
        cnt[q] = len(sel)
    return torch.from_numpy(out.view(np.int32).reshape(nq, k, 4)), torch.from_numpy(cnt)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from coltt_b200.dist import HIT_DTYPE, ShardedSearch, gpu_of, unpack_hits
    from oracle import oracle as orc
    from tests.util import QUERY_SEED, normal, sparse_ids
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, d, k = 4000, 64, 10
    ids, vecs = sparse_ids(n), normal(n, d)
    mine = gpu_of(ids, world) == rank
    st = orc.FlatStore(d, orc.COSINE, orc.Q_F16)
    st.upsert(ids[mine], vecs[mine])
    full = orc.FlatStore(d, orc.COSINE, orc.Q_F16)
    full.upsert(ids, vecs)
    qs = normal(6, d, QUERY_SEED)

    def local_search(queries, kk, mode):
        hits = np.zeros((len(queries), kk), dtype=HIT_DTYPE)
        cnt = np.zeros(len(queries), dtype=np.int32)
        for j, q in enumerate(queries):
            i_, s_ = st.search_total_order(q, kk, select_mode=mode)
            hits["id"][j, : len(i_)] = i_
            hits["score"][j, : len(i_)] = s_
            cnt[j] = len(i_)
        return torch.from_numpy(hits.view(np.int32).reshape(len(queries), kk, 4)), torch.from_numpy(cnt)

    ss = ShardedSearch(local_search, _numpy_merge)
    ok = True
    for mode in (0, 1):
        hits, cnt = ss.search(qs, k, mode)
        gi, gs, gc = unpack_hits(hits, cnt)
        for j, q in enumerate(qs):
            wi, ws = full.search_total_order(q, k, select_mode=mode)
            ok &= bool(np.array_equal(gi[j, : gc[j]], wi) and gs[j, : gc[j]].tobytes() == ws.tobytes())
    ret[rank] = ok and int(mine.sum()) > 0
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_search_equals_single_store():
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 1000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
