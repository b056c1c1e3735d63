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
    from coltt_b200.dist import ShardedSearch, cuda_callables, gpu_of, unpack_hits
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
        sp.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_GPU_OK" if int(t.item()) == 1 else "DIST_GPU_FAIL", "world", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
