"""Index-arithmetic check of the experimental RING staging of hnsw_search_kernel (csrc/hnsw.cu, COLTT_HNSW_RING), on
the CPU: the per-warp chunk ring is transcribed into Python — copies land at issue time (the earliest they could, so a
refill issued before the reads of that stage would corrupt the result), stage barriers count issues and waits — and the
accumulation must equal the whole-row computation in the same AVX-lane order, for ragged dims, chunk sizes and depths."""
import numpy as np
import pytest

f32 = np.float32


def whole_row(q, row, dim):
    """4 lanes per row, 2 AVX-lane chains each, then the reduction tree and the scalar tail (the non-RING path)."""
    full8 = dim // 8 * 8
    acc = [f32(0)] * 8
    for e in range(0, full8, 8):
        for l in range(8):
            acc[l] = f32(acc[l] + f32(q[e + l] * row[e + l]))
    t = [f32(acc[2 * h] + acc[2 * h + 1]) for h in range(4)]
    u = [f32(t[0] + t[1]), f32(t[2] + t[3])]
    tot = f32(u[0] + u[1])
    for d in range(full8, dim):
        tot = f32(tot + f32(q[d] * row[d]))
    return tot


def ring_row_group(q, rows, dim, row_stride, CB, S):
    """One warp, `len(rows)` <= 8 rows: returns the per-row totals computed through the ring."""
    nrows = len(rows)
    CS = (CB + 127) // 128 * 128 + 32
    full8 = dim // 8 * 8
    CBE = CB // 4
    n_chunks = (row_stride + CB - 1) // CB
    ring = np.full((S, 8, CS // 4), np.nan, dtype=np.float32)          # poisoned
    src = np.zeros((nrows, row_stride // 4), np.float32)
    for r in range(nrows):
        src[r, :dim] = rows[r]
    pending = [0] * S        # copies issued and not yet waited for, per stage
    read_done = [True] * S   # the stage's previous contents were consumed

    def issue(c):
        st, off = c % S, c * CB
        nbytes = min(CB, row_stride - off)
        assert read_done[st], "refill before the previous chunk of this stage was read"
        assert off % 16 == 0 and nbytes % 16 == 0 and (CS * 4) % 16 == 0
        for r in range(nrows):
            ring[st, r, : nbytes // 4] = src[r, off // 4: off // 4 + nbytes // 4]
        pending[st] += 1
        read_done[st] = False

    for c in range(min(S, n_chunks)):
        issue(c)
    a = np.zeros((nrows, 8), np.float32)
    tot = [None] * nrows
    for c in range(n_chunks):
        st = c % S
        assert pending[st] == 1, "wait without a matching issue"
        pending[st] -= 1
        e0 = c * CBE
        e1 = min(e0 + CBE, full8)
        for r in range(nrows):
            for e in range(e0, e1, 8):
                for h in range(4):
                    rv = ring[st, r, e - e0 + 2 * h: e - e0 + 2 * h + 2]
                    a[r, 2 * h] = f32(a[r, 2 * h] + f32(q[e + 2 * h] * rv[0]))
                    a[r, 2 * h + 1] = f32(a[r, 2 * h + 1] + f32(q[e + 2 * h + 1] * rv[1]))
            if c == n_chunks - 1:
                t = [f32(a[r, 2 * h] + a[r, 2 * h + 1]) for h in range(4)]
                x = f32(f32(t[0] + t[1]) + f32(t[2] + t[3]))
                for d in range(full8, dim):
                    x = f32(x + f32(q[d] * ring[st, r, d - e0]))
                tot[r] = x
        read_done[st] = True
        if c + S < n_chunks:
            issue(c + S)
    assert all(p == 0 for p in pending)
    return tot


@pytest.mark.parametrize("dim", [8, 9, 17, 100, 130, 257, 768, 770, 775])
def test_ring_staging_reads_every_element_once_in_reference_order(dim):
    rng = np.random.default_rng(dim)
    row_stride = (dim * 4 + 15) // 16 * 16
    q = rng.standard_normal(dim).astype(np.float32)
    for CB in (32, 64, 512, 768, 4096):
        cb = min(CB, (row_stride + 31) // 32 * 32)          # the host clamps the chunk to the (32-byte padded) row
        for S in (2, 3, 4):
            for nrows in (1, 5, 8):
                rows = rng.standard_normal((nrows, dim)).astype(np.float32)
                got = ring_row_group(q, rows, dim, row_stride, cb, S)
                want = [whole_row(q, rows[r], dim) for r in range(nrows)]
                assert [x.tobytes() for x in got] == [x.tobytes() for x in want], (dim, CB, S, nrows)
