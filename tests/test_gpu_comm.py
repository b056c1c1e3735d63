"""The sharded-search entry points of the C-ABI on ONE GPU (world size 1: the local search writes into the rank's send buffer,
no all-gather, K5 merges the single list).  The driver's 1-GPU box exercises csrc/comm.cu through these; the multi-GPU cases
live in tests/test_gpu_dist.py."""
import ctypes as C

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


def test_world_one_sharded_search_equals_the_store_search(cb, oracle):
    import torch
    from coltt_b200 import _lib
    from coltt_b200.dist import Comm
    L = _lib.lib()
    n, d, k, nq = 50_000, 256, 10, 40
    ids, vecs = sparse_ids(n), normal(n, d)
    qs = normal(nq, d, QUERY_SEED)
    comm = Comm.from_torch_distributed(0)          # no process group: coltt_b200_init([0], 1)
    r, w, dev = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    _lib.check(L.coltt_b200_comm_info(comm._h, C.byref(r), C.byref(w), C.byref(dev)))
    assert (r.value, w.value, dev.value) == (0, 1, 0)
    for quant, math in ((cb.Quantization_BF16, cb.MATH_FAST), (cb.Quantization_None, cb.MATH_FAST), (cb.Quantization_F8, cb.MATH_EXACT)):
        sp = cb.VectorSpace("s", cb.Metadata(d, cb.Distance_Cosine, quant))
        sp.ChangedVertices(ids, vecs)
        st = oracle.FlatStore(d, oracle.COSINE, quant)
        st.upsert(ids, vecs)
        for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
            gi, gs, gc = comm.search(sp, qs, k, mode, math)                    # host buffers
            si, ss, sc_ = sp.BatchVertexSearch(qs, k, select_mode=mode, math_mode=math)
            assert np.array_equal(gi, si) and gs.tobytes() == ss.tobytes() and np.array_equal(gc, sc_)
            for j in (0, 17, 39):
                wi, ws = st.search_total_order(qs[j], k, select_mode=mode)
                assert_same_hits(gi[j, :gc[j]], gs[j, :gc[j]], wi, ws, f"quant={quant} mode={mode} q{j}")
            # device-resident form on a caller stream
            q = torch.from_numpy(qs).cuda()
            out = torch.empty((nq, k, 4), dtype=torch.int32, device="cuda")
            cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
            s = torch.cuda.Stream()
            torch.cuda.synchronize()
            for _ in range(3):                                                   # third call replays the captured graph
                comm.search_dev(sp, q.data_ptr(), nq, k, mode, math, out.data_ptr(), cnt.data_ptr(), s.cuda_stream)
            s.synchronize()
            h = out.cpu().numpy().view(np.uint8).reshape(nq, k, 16)
            di, ds = h[..., :8].copy().view(np.uint64)[..., 0], h[..., 8:12].copy().view(np.float32)[..., 0]
            assert np.array_equal(di, si) and ds.tobytes() == ss.tobytes()
        sp.close()
    # HNSW through the communicator (one sub-graph = the whole index), fp32 and PQ walks
    nh, dh = 6000, 64
    hv, hid = normal(nh, dh, 5), sparse_ids(nh, 9)
    hq = normal(12, dh, QUERY_SEED + 1)
    h = cb.Hnsw.Build(hid, hv, metric=cb.Distance_Cosine, m=16, ef=64)
    a = comm.hnsw_search(h, hq, k, 64)
    b = h.BatchSearch(hq, k, 64)
    assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()
    h.TrainPQ(256, 16, 4096)
    a = comm.hnsw_search(h, hq, k, 64, pq=True)
    b = h.BatchSearchPQ(hq, k, 64, rerank=True)
    assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()
    h.close()
    comm.close()
    # dimension / argument errors come back as codes, not crashes
    with pytest.raises(cb.ColttError):
        comm2 = Comm.from_torch_distributed(0)
        try:
            sp = cb.VectorSpace("e", cb.Metadata(8, cb.Distance_Cosine, cb.Quantization_None))
            comm2.search(sp, np.zeros((1, 8), np.float32), 0, cb.SELECT_NEAREST, cb.MATH_EXACT)   # k = 0
        finally:
            comm2.close()
