"""Micro-batcher for the single-query RPC path.

The reference serves one vector per Search RPC, each in its own goroutine (edge/edge.go:610-690, edge.proto:119-126).
The GPU path wants Q >= 256 queries per launch (the tcgen05 filter costs the same for 1 and for 256
queries), so concurrent single-query callers are coalesced here: a caller enqueues its query and blocks on a future;
one flusher thread takes up to `max_batch` queued queries of equal topK — as soon as `max_batch` are waiting, or
`max_wait_ms` after the first one arrived — runs ONE batched search and hands every caller its row.  Results are the
batched call's rows, which are bit-identical to single-query calls (tests/test_gpu_fast.py).  The Go twin is
`bridge/go/colttb200/batcher.go`."""
import threading
import time
from concurrent.futures import Future
from typing import Callable, List, Tuple

import numpy as np


class MicroBatcher:
    def __init__(self, batch_search: Callable[[np.ndarray, int], Tuple[np.ndarray, np.ndarray, np.ndarray]], dim: int,
                 max_batch: int = 256, max_wait_ms: float = 0.2):
        """`batch_search(queries [nq, dim] f32, topK) -> (ids [nq, k], scores [nq, k], counts [nq])`, e.g.
        `VectorSpace.BatchVertexSearch`."""
        self._search, self.dim, self.max_batch, self.max_wait = batch_search, dim, max_batch, max_wait_ms / 1e3
        self._cv = threading.Condition()
        self._queue: List[Tuple[np.ndarray, int, Future, float]] = []
        self._closed = False
        self.batches = 0           # statistics: batched searches issued / queries served
        self.served = 0
        self._thread = threading.Thread(target=self._run, name="coltt-b200-batcher", daemon=True)
        self._thread.start()

    def submit(self, target, topK: int) -> Future:
        q = np.ascontiguousarray(target, dtype=np.float32).reshape(-1)
        if q.size != self.dim:
            raise ValueError("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]" % (self.dim, q.size))
        if topK <= 0:
            raise ValueError("topK must be positive")
        f: Future = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("batcher is closed")
            self._queue.append((q, int(topK), f, time.perf_counter()))
            self._cv.notify_all()
        return f

    def VertexSearch(self, target, topK: int):
        """Blocking single-query call: (ids [count], scores [count]) of this query, as VertexSearch returns them."""
        return self.submit(target, topK).result()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join()

    def _take(self):
        """Called with the lock held: the next batch (same topK as the oldest request), or None to keep waiting."""
        if not self._queue:
            return None
        k0 = self._queue[0][1]
        same = [i for i, it in enumerate(self._queue) if it[1] == k0][: self.max_batch]
        full = len(same) >= self.max_batch
        due = time.perf_counter() - self._queue[0][3] >= self.max_wait
        if not (full or due or self._closed):
            return None
        batch = [self._queue[i] for i in same]
        for i in reversed(same):
            del self._queue[i]
        return batch

    def _run(self):
        while True:
            with self._cv:
                batch = self._take()
                while batch is None:
                    if self._closed and not self._queue:
                        return
                    wait = None
                    if self._queue:
                        wait = max(0.0, self.max_wait - (time.perf_counter() - self._queue[0][3]))
                    self._cv.wait(timeout=wait)
                    batch = self._take()
            qs = np.stack([b[0] for b in batch])
            k = batch[0][1]
            try:
                ids, sc, cnt = self._search(qs, k)
                for j, b in enumerate(batch):
                    c = int(cnt[j])
                    b[2].set_result((np.array(ids[j, :c]), np.array(sc[j, :c])))
            except Exception as e:  # every waiting caller sees the error, like the RPC's failFn
                for b in batch:
                    if not b[2].done():
                        b[2].set_exception(e)
            self.batches += 1
            self.served += len(batch)
