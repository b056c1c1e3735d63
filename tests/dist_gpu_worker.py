"""torchrun worker for tests/test_gpu_dist.py: one rank per GPU, NCCL, rows partitioned by ShardVertex."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import coltt_b200 as cb
    from coltt_b200.dist import Comm, ShardedSearch, cuda_callables, gpu_of, unpack_hits
    from oracle import oracle as orc
    from tests.util import QUERY_SEED, normal, sparse_ids
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, d, k = 60_000, 256, 10
    ids, vecs = sparse_ids(n), normal(n, d)
    mine = gpu_of(ids, world) == rank
    ok = True
    for quant, math in ((cb.Quantization_BF16, cb.MATH_FAST), (cb.Quantization_None, cb.MATH_EXACT)):
        sp = cb.VectorSpace("s", cb.Metadata(d, cb.Distance_Cosine, quant), device=local)
        sp.ChangedVertices(ids[mine], vecs[mine])
        ls, mg = cuda_callables(sp, local, math_mode=math)
        ss = ShardedSearch(ls, mg)
        qs = normal(40, d, QUERY_SEED)
        for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
            hits, cnt = ss.search(qs, k, mode)
            torch.cuda.synchronize()
            gi, gs, gc = unpack_hits(hits, cnt)
            if rank == 0:
                full = orc.FlatStore(d, orc.COSINE, quant)
                full.upsert(ids, vecs)
                for j in range(8):
                    wi, ws = full.search_total_order(qs[j], k, select_mode=mode)
                    good = np.array_equal(gi[j, : gc[j]], wi) and gs[j, : gc[j]].tobytes() == ws.tobytes()
                    if not good:
                        print("MISMATCH", quant, mode, j, gi[j], wi, flush=True)
                    ok &= bool(good)
        # the same search behind the C-ABI (csrc/comm.cu): local search + ncclAllGather + merge inside the library
        comm = Comm.from_torch_distributed(local)
        for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
            gi, gs, gc = comm.search(sp, qs, k, mode, math)
            if rank == 0:
                for j in range(8):
                    wi, ws = full.search_total_order(qs[j], k, select_mode=mode)
                    good = np.array_equal(gi[j, : gc[j]], wi) and gs[j, : gc[j]].tobytes() == ws.tobytes()
                    if not good:
                        print("MISMATCH (C-ABI comm)", quant, mode, j, gi[j], wi, flush=True)
                    ok &= bool(good)
        dist.barrier()      # a peer may still be merging this rank's last list out of its exchange buffer (see coltt_b200_comm_destroy)
        comm.close()
        sp.close()
    # HNSW shards: one independent sub-graph per GPU, same all-gather + merge; the merged answer must equal the
    # host-side merge of the oracle's walks over the same per-shard graphs
    nh, dh = 8000, 64
    hv, hid = normal(nh, dh, 123), sparse_ids(nh, 77)
    hq = normal(16, dh, QUERY_SEED + 5)
    hm = gpu_of(hid, world) == rank
    sub = cb.Hnsw.Build(hid[hm], hv[hm], metric=cb.Distance_Cosine, m=16, ef=64, device=local)
    comm = Comm.from_torch_distributed(local)
    gi, gs, gc = comm.hnsw_search(sub, hq, k, 64)
    oh = orc.Hnsw.load(sub.Commit())
    oh.set_ef(64)
    loc = np.zeros((len(hq), k, 2), np.float64)
    for j in range(len(hq)):
        wi, ws = oh.search(hq[j], k)
        loc[j, : len(wi), 0] = wi.astype(np.float64)      # ids < 2^62 are not exact in f64: compare through the same cast below
        loc[j, : len(wi), 1] = ws
        loc[j, len(wi):, 1] = np.inf
    tl = torch.from_numpy(loc).cuda()
    parts = [torch.zeros_like(tl) for _ in range(world)]
    dist.all_gather(parts, tl)
    if rank == 0:
        allp = np.concatenate([p_.cpu().numpy() for p_ in parts], axis=1)
        for j in range(len(hq)):
            order = np.lexsort((allp[j, :, 0], allp[j, :, 1]))[:k]
            good = np.array_equal(gi[j, :k].astype(np.float64), allp[j, order, 0]) and np.array_equal(gs[j, :k].astype(np.float64), allp[j, order, 1])
            if not good:
                print("MISMATCH (sharded hnsw)", j, gi[j], allp[j, order, 0], flush=True)
            ok &= bool(good)
    exchange = comm.exchange
    dist.barrier()
    comm.close()
    sub.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_GPU_OK" if int(t.item()) == 1 else "DIST_GPU_FAIL", "world", world, "exchange", exchange, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
