"""Multi-GPU sharded search on real GPUs (skipped with fewer than 2): rows partitioned by ShardVertex,
local tcgen05/exact search per rank, one NCCL all-gather of per-shard top-k, K5 merge — must equal a
single store over all rows (the oracle)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["peer-memory", "nccl"])
def test_sharded_search_two_or_more_gpus(exchange):
    """Both exchange steps of csrc/comm.cu — the merge kernel reading the peers' lists over NVLink (default where every pair of
    ranks has peer access) and ncclAllGather + merge (COLTT_P2P=0) — against the oracle over the union of the shards."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    env = dict(os.environ, COLTT_P2P="1" if exchange == "peer-memory" else "0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert "DIST_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    used = out.stdout.split("exchange", 1)[1].split()[0]
    if exchange == "nccl":
        assert used == "nccl", out.stdout[-500:]
    elif used != "peer-memory":
        pytest.skip("this box has no peer access between the GPUs: the ranks voted for the NCCL exchange")


def test_single_process_multi_gpu_through_the_c_abi():
    """coltt_b200_init(device_ids, n) + coltt_b200_sharded_search_all: ONE process drives every GPU (the Go host's shape,
    INTEGRATION.md) — ncclCommInitAll, one library thread per rank, result == the oracle over the union of the shards."""
    import ctypes as C
    import numpy as np
    import torch
    import coltt_b200 as cb
    from coltt_b200 import _lib
    from coltt_b200.dist import gpu_of
    from oracle import oracle as orc
    from tests.util import QUERY_SEED, normal, sparse_ids
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    g = 2 if g < 4 else 4
    L = _lib.lib()
    n, d, k = 50_000, 256, 10
    ids, vecs = sparse_ids(n), normal(n, d)
    owner = gpu_of(ids, g)
    shards = []
    for r in range(g):
        sp = cb.VectorSpace("s", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_BF16), device=r)
        sp.ChangedVertices(ids[owner == r], vecs[owner == r])
        shards.append(sp)
    devs = (C.c_int * g)(*range(g))
    comms = (C.c_void_p * g)()
    _lib.check(L.coltt_b200_init(devs, g, comms))
    sh = (C.c_void_p * g)(*[s._h for s in shards])
    qs = normal(40, d, QUERY_SEED)
    full = orc.FlatStore(d, orc.COSINE, orc.Q_BF16)
    full.upsert(ids, vecs)
    for mode in (cb.SELECT_NEAREST, cb.SELECT_COMPAT):
        for mm in (cb.MATH_FAST, cb.MATH_EXACT):
            oi, osc, oc = np.zeros((40, k), np.uint64), np.zeros((40, k), np.float32), np.zeros(40, np.int32)
            _lib.check(L.coltt_b200_sharded_search_all(comms, sh, g, qs.ctypes.data_as(C.POINTER(C.c_float)), 40, k, mode, mm,
                                                       oi.ctypes.data_as(C.POINTER(C.c_uint64)), osc.ctypes.data_as(C.POINTER(C.c_float)),
                                                       oc.ctypes.data_as(C.POINTER(C.c_int32))))
            for j in range(8):
                wi, ws = full.search_total_order(qs[j], k, select_mode=mode)
                assert np.array_equal(oi[j, : oc[j]], wi) and osc[j, : oc[j]].tobytes() == ws.tobytes(), (mode, mm, j)
    L.coltt_b200_shutdown()
    for s in shards:
        s.close()


_TIMEOUT_SCRIPT = r"""
import ctypes as C, sys
import numpy as np
import coltt_b200 as cb
from coltt_b200 import _lib
from tests.util import normal, sparse_ids
L = _lib.lib()
g, n, d, k, nq = 2, 4000, 64, 5, 8
ids, vecs = sparse_ids(n), normal(n, d)
shards = []
for r in range(g):
    sp = cb.VectorSpace("t", cb.Metadata(d, cb.Distance_Cosine, cb.Quantization_None), device=r)
    sp.ChangedVertices(ids[r::g], vecs[r::g])
    shards.append(sp)
comms = (C.c_void_p * g)()
_lib.check(L.coltt_b200_init((C.c_int * g)(0, 1), g, comms))
sh = (C.c_void_p * g)(*[s._h for s in shards])
qs = normal(nq, d, 9)
oi, osc, oc = np.zeros((nq, k), np.uint64), np.zeros((nq, k), np.float32), np.zeros(nq, np.int32)
args = (qs.ctypes.data_as(C.POINTER(C.c_float)), nq, k, cb.SELECT_NEAREST, cb.MATH_EXACT, oi.ctypes.data_as(C.POINTER(C.c_uint64)),
        osc.ctypes.data_as(C.POINTER(C.c_float)), oc.ctypes.data_as(C.POINTER(C.c_int32)))
_lib.check(L.coltt_b200_sharded_search_all(comms, sh, g, *args))          # a proper collective: sets the exchange up
if L.coltt_b200_comm_exchange_mode(comms[0]) != 1:
    print("NO_PEER_ACCESS")
    sys.exit(0)
rc = L.coltt_b200_sharded_search(comms[0], sh[0], *args)                  # rank 1 never arrives
msg = (L.coltt_b200_last_error() or b"").decode()
rc2 = L.coltt_b200_sharded_search(comms[0], sh[0], *args)                 # the communicator stays failed
print("RC", rc, rc2, "|", msg)
L.coltt_b200_shutdown()
"""


def test_a_peer_that_never_arrives_is_an_error_not_a_hang():
    """The peer-memory exchange waits for the peers' flags with a deadline (COLTT_P2P_TIMEOUT_MS): a rank whose peer does not
    join the collective gets COLTT_ERR_CUDA with a message naming the step, within the deadline, and the communicator then
    refuses further searches (the ranks' sequence numbers no longer agree)."""
    import time
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, COLTT_P2P="1", COLTT_P2P_TIMEOUT_MS="400", PYTHONPATH=ROOT)
    t0 = time.time()
    out = subprocess.run([sys.executable, "-c", _TIMEOUT_SCRIPT], capture_output=True, text=True, timeout=180, cwd=ROOT, env=env)
    if "NO_PEER_ACCESS" in out.stdout:
        pytest.skip("no peer access between the GPUs of this box")
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RC")]
    assert line, out.stdout[-1500:] + out.stderr[-1500:]
    rc, rc2 = int(line[0].split()[1]), int(line[0].split()[2])
    assert rc == -2 and rc2 == -2 and "did not publish its results within the exchange timeout" in line[0], line[0]
    assert time.time() - t0 < 120
