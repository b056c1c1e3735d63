"""Lane-level check of WarpSorted (csrc/hnsw.cu), on the CPU: the struct's methods transcribed statement by statement
into Python over explicit 32-lane state (shfl_up / shfl_down / shfl / ballot / any spelled out), driven by random operation
sequences and compared with the plain sorted-list semantics the queue model (test_hnsw_fastpath_model.py) assumes:
blocked layout (lane L owns entries [L*R, L*R+R)), insert before equal priorities, eviction past ef, removal by slot,
expanded flags, the H1/H2/H3 probes.  Covers the lane-boundary cases real data rarely reaches."""
import math

import numpy as np
import pytest

INF = float("inf")
NOSLOT = 0xFFFFFFFF
EXP = 0x80000000


class LaneWS:
    def __init__(self, R):
        self.R = R
        self.p = [[INF] * R for _ in range(32)]
        self.s = [[NOSLOT] * R for _ in range(32)]
        self.n = 0
        self.has_tie = False

    # --- warp collectives over per-lane values
    @staticmethod
    def shfl_up(v):      # lane L reads lane L-1; lane 0 keeps its own value
        return [v[0]] + v[:-1]

    @staticmethod
    def shfl_down(v):    # lane L reads lane L+1; lane 31 keeps its own value
        return v[1:] + [v[-1]]

    def from_top(self, back):
        R, lim = self.R, self.n - back
        v = []
        for lane in range(32):
            x = self.p[lane][0]
            for r in range(1, R):
                if lane * R + r < lim:
                    x = self.p[lane][r]
            v.append(x)
        return v[((lim - 1) // R) & 31]

    def insert(self, d, slot, ef, check=True):
        R = self.R
        cnt = [sum(1 for r in range(R) if self.p[l][r] < d) for l in range(32)]
        eq = [any(self.p[l][r] == d for r in range(R)) for l in range(32)]
        up_p = self.shfl_up([self.p[l][R - 1] for l in range(32)])
        up_s = self.shfl_up([self.s[l][R - 1] for l in range(32)])
        below = [cnt[l] == R for l in range(32)]
        if d != d:
            return False
        if any(eq):
            self.has_tie = True
        if check and self.has_tie and self.n >= ef:
            m1 = self.from_top(0)
            if d == m1:
                return False
            if d < m1 and self.n >= 2 and self.from_top(1) == m1:
                return False
        full = sum(below)
        for lane in range(32):
            p, s = self.p[lane], self.s[lane]
            if lane > full:
                for r in range(R - 1, 0, -1):
                    p[r], s[r] = p[r - 1], s[r - 1]
                p[0], s[0] = up_p[lane], up_s[lane]
            elif lane == full:
                c = cnt[lane]
                for r in range(R - 1, 0, -1):
                    if r > c:
                        p[r], s[r] = p[r - 1], s[r - 1]
                    elif r == c:
                        p[r], s[r] = d, slot
                if c == 0:
                    p[0], s[0] = d, slot
        self.n += 1
        if self.n > ef:
            self.n = ef
            for lane in range(32):
                for r in range(R):
                    if lane * R + r >= ef:
                        self.p[lane][r], self.s[lane][r] = INF, NOSLOT
        return True

    def remove(self, slot):
        R = self.R
        hit = []
        for lane in range(32):
            h = R
            for r in range(R):
                if (self.s[lane][r] & ~EXP) == slot and lane * R + r < self.n:
                    h = r
            hit.append(h)
        have = [h < R for h in hit]
        dn_p = self.shfl_down([self.p[l][0] for l in range(32)])
        dn_s = self.shfl_down([self.s[l][0] for l in range(32)])
        if not any(have):
            return
        dn_p[31], dn_s[31] = INF, NOSLOT
        at = have.index(True)
        for lane in range(32):
            if lane >= at:
                p, s = self.p[lane], self.s[lane]
                for r in range(R - 1):
                    if lane > at or r >= hit[lane]:
                        p[r], s[r] = p[r + 1], s[r + 1]
                p[R - 1], s[R - 1] = dn_p[lane], dn_s[lane]
        self.n -= 1

    def find(self):
        R = self.R
        found, vp, vs = [], [], []
        for lane in range(32):
            f, a, b = False, 0.0, 0
            for r in range(R):
                c = (not f) and not (self.s[lane][r] & EXP) and lane * R + r < self.n
                if c:
                    a, b = self.p[lane][r], self.s[lane][r]
                f = f or c
            found.append(f); vp.append(a); vs.append(b)
        if not any(found):
            return None
        src = found.index(True)
        return vp[src], vs[src]

    def several(self, cp):
        R = self.R
        c = [sum(1 for r in range(R) if not (self.s[l][r] & EXP) and l * R + r < self.n and self.p[l][r] == cp) for l in range(32)]
        return sum(1 for x in c if x > 0) > 1 or any(x > 1 for x in c)

    def mark(self, slot):
        for lane in range(32):
            for r in range(self.R):
                if self.s[lane][r] == slot and lane * self.R + r < self.n:
                    self.s[lane][r] |= EXP

    def head_ties(self, k):
        R = self.R
        lim = min(k + 1, self.n)
        nxt0 = self.shfl_down([self.p[l][0] for l in range(32)])
        t = False
        for lane in range(32):
            for r in range(R):
                nx = self.p[lane][r + 1] if r + 1 < R else nxt0[lane]
                t = t or (lane * R + r + 1 < lim and self.p[lane][r] == nx)
        return t

    def entries(self):
        out = []
        for i in range(self.n):
            lane, r = divmod(i, self.R)
            out.append((self.p[lane][r], self.s[lane][r] & ~EXP, bool(self.s[lane][r] & EXP)))
        return out

    def padding_ok(self):
        return all(self.p[i // self.R][i % self.R] == INF and self.s[i // self.R][i % self.R] == NOSLOT
                   for i in range(self.n, 32 * self.R))


class ListWS:
    """The semantics the queue model assumes."""
    def __init__(self):
        self.e, self.has_tie = [], False

    def insert(self, d, slot, ef, check=True):
        if d != d:
            return False
        if any(x[0] == d for x in self.e):
            self.has_tie = True
        if check and self.has_tie and len(self.e) >= ef:
            m1 = self.e[-1][0]
            if d == m1 or (d < m1 and len(self.e) >= 2 and self.e[-2][0] == m1):
                return False
        self.e.insert(sum(1 for x in self.e if x[0] < d), [d, slot, False])
        if len(self.e) > ef:
            self.e.pop()
        return True


@pytest.mark.parametrize("R", [2, 4, 8])
def test_warpsorted_lane_code_equals_sorted_list_semantics(R):
    rng = np.random.default_rng(R)
    for trial in range(250):
        ef = int(rng.choice([1, 2, 3, R, R + 1, 2 * R, 31, 32, 33, 16 * R, 32 * R - 1, 32 * R]))
        ef = min(ef, 32 * R)
        levels = int(rng.choice([3, 10, 1000]))
        a, b = LaneWS(R), ListWS()
        slot = 0
        for step in range(int(rng.integers(1, 4 * ef + 20))):
            op = rng.random()
            if op < 0.7 or not b.e:
                d = float(rng.integers(0, levels)) if rng.random() > 0.01 else math.nan
                slot += 1
                check = rng.random() < 0.8
                if not check and len(b.e) >= ef:
                    # the unchecked insert only ever follows a remove (n == ef - 1): reproduce that pairing
                    victim = b.e[int(rng.integers(0, len(b.e)))][1]
                    a.remove(victim)
                    b.e = [x for x in b.e if x[1] != victim]
                ra, rb = a.insert(d, slot, ef, check), b.insert(d, slot, ef, check)
                assert ra == rb, (trial, step, "insert result")
            elif op < 0.8:
                victim = b.e[int(rng.integers(0, len(b.e)))][1] if rng.random() < 0.9 else 123456789
                a.remove(victim)
                b.e = [x for x in b.e if x[1] != victim]
            else:
                got = a.find()
                idx = next((i for i, x in enumerate(b.e) if not x[2]), None)
                if idx is None:
                    assert got is None
                else:
                    assert got == (b.e[idx][0], b.e[idx][1])
                    assert a.several(got[0]) == (sum(1 for x in b.e if not x[2] and x[0] == got[0]) > 1)
                    a.mark(got[1])
                    b.e[idx][2] = True
            assert a.n == len(b.e) and a.has_tie == b.has_tie
            assert a.entries() == [tuple(x) for x in b.e], (trial, step)
            assert a.padding_ok()
            if b.e:
                assert a.from_top(0) == b.e[-1][0]
                if len(b.e) >= 2:
                    assert a.from_top(1) == b.e[-2][0]
                k = int(rng.integers(1, ef + 1))
                lim = min(k + 1, len(b.e))
                assert a.head_ties(k) == any(b.e[i][0] == b.e[i + 1][0] for i in range(lim - 1))


class HoleHeap:
    """SmemHeap of csrc/hnsw.cu: up / down move a hole instead of swapping."""
    def __init__(self, is_max):
        self.e, self.is_max = [], is_max

    def less(self, a, b):
        return a > b if self.is_max else a < b

    def push(self, p, s):
        self.e.append(None)
        j = len(self.e) - 1
        while j > 0:
            i = (j - 1) // 2
            if not self.less(p, self.e[i][0]):
                break
            self.e[j] = self.e[i]
            j = i
        self.e[j] = (p, s)

    def pop(self):
        m = len(self.e) - 1
        t = self.e[0]
        if m > 0:
            x = self.e[m]
            i = 0
            while True:
                j1 = 2 * i + 1
                if j1 >= m:
                    break
                c, j = self.e[j1], j1
                if j1 + 1 < m and self.less(self.e[j1 + 1][0], c[0]):
                    c, j = self.e[j1 + 1], j1 + 1
                if not self.less(c[0], x[0]):
                    break
                self.e[i] = c
                i = j
            self.e[i] = x
        self.e.pop()
        return t


@pytest.mark.parametrize("is_max", [False, True])
def test_hole_moving_sifts_equal_go_container_heap(is_max):
    """The array after every push / pop, and every popped element, equal Go's swap-based heap — ties and NaN included."""
    from tests.test_hnsw_fastpath_model import GoHeap
    rng = np.random.default_rng(int(is_max))
    for trial in range(300):
        g, h = GoHeap(is_max), HoleHeap(is_max)
        levels = int(rng.choice([2, 5, 1000]))
        for step in range(int(rng.integers(1, 200))):
            if rng.random() < 0.6 or not g.a:
                p = float(rng.integers(0, levels)) if rng.random() > 0.02 else math.nan
                g.push(p, step)
                h.push(p, step)
            else:
                assert repr(g.pop()) == repr(h.pop())
            assert repr(g.a) == repr(h.e), (trial, step)
