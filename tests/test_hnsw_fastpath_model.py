"""Design check of K4's queue handling (csrc/hnsw.cu), on the CPU: a step-by-step Python model of what warp 0 does —
the ef-entry result set as a sorted array with `expanded` flags, the operation log, and the literal Go heaps rebuilt
lazily from that log only at the three hazards (H1 candidate pop among equal minima, H2 eviction with equal maxima,
H3 equal priorities among the first k+1 results) — must walk exactly like a literal restatement of
searchLevel (core/vectorindex/hnsw.go:345-389) over Go's container/heap, on random graphs whose distances are small
integers (ties everywhere).  Same results, same order, same evaluation and expansion counts."""
import numpy as np
import pytest


# ---- Go container/heap, keyed on priority only (core/vectorindex/priority_queue.go:160-199) -------------------
class GoHeap:
    def __init__(self, is_max):
        self.a, self.is_max = [], is_max

    def less(self, x, y):
        return x[0] > y[0] if self.is_max else x[0] < y[0]

    def push(self, p, s):
        a = self.a
        a.append((p, s))
        j = len(a) - 1
        while j > 0:
            i = (j - 1) // 2
            if i == j or not self.less(a[j], a[i]):
                break
            a[i], a[j] = a[j], a[i]
            j = i

    def pop(self):
        a = self.a
        n = len(a) - 1
        a[0], a[n] = a[n], a[0]
        i = 0
        while True:
            j1 = 2 * i + 1
            if j1 >= n:
                break
            j = j1
            if j1 + 1 < n and self.less(a[j1 + 1], a[j1]):
                j = j1 + 1
            if not self.less(a[j], a[i]):
                break
            a[i], a[j] = a[j], a[i]
            i = j
        return a.pop()

    def __len__(self):
        return len(self.a)


def literal_search(nbrs, dist, ep, ef, k):
    """searchLevel + selectNeighbors + the back-to-front fill of Hnsw.Search, literally."""
    cand, res = GoHeap(False), GoHeap(True)
    cand.push(dist[ep], ep)
    res.push(dist[ep], ep)
    visited = {ep}
    evals, exps = 1, 0
    while len(cand):
        cp, cs = cand.pop()
        lb = res.a[0][0]
        if cp > lb:
            break
        exps += 1
        for s in nbrs[cs]:
            if s in visited:
                continue
            visited.add(s)
            d = dist[s]
            evals += 1
            if d < lb or len(res) < ef:
                cand.push(d, s)
                res.push(d, s)
                if len(res) > ef:
                    res.pop()
    while len(res) > k:
        res.pop()
    out = [None] * len(res)
    for i in range(len(res) - 1, -1, -1):
        out[i] = res.pop()
    return out, evals, exps


def model_search(nbrs, dist, ep, ef, k, stats):
    """What hnsw_search_kernel<METRIC, R > 0> does with its queues (no NaN in this model)."""
    ws = []                      # sorted ascending by priority: [p, slot, expanded]
    has_tie = False
    log = []                     # (p, slot) pushes and None pop markers
    cand, res = GoHeap(False), GoHeap(True)
    pos = {"cand": 0, "res": 0}
    haz_set, haz_x = False, 0.0

    def replay(which):
        h = cand if which == "cand" else res
        for e in log[pos[which]:]:
            if e is None:
                if which == "cand":
                    h.pop()
            else:
                h.push(*e)
                if which == "res" and len(h) > ef:
                    h.pop()
        pos[which] = len(log)

    def insert(d, s, check=True):
        nonlocal has_tie
        if d != d:
            return False                              # NaN: the walk continues on the literal heaps for good
        if any(e[0] == d for e in ws):
            has_tie = True
        if check and has_tie and len(ws) >= ef:
            m1 = ws[-1][0]
            if d == m1 or (d < m1 and len(ws) >= 2 and ws[-2][0] == m1):
                return False
        at = sum(1 for e in ws if e[0] < d)
        ws.insert(at, [d, s, False])
        if len(ws) > ef:
            ws.pop()
        return True

    def finish_literal(out_of, lb, rest):
        """go_literal(): both heaps brought up to date, then the reference's loop — first the rest of the expansion
        that met the NaN (same frozen lowerBound), then the remaining walk."""
        nonlocal evals, exps
        replay("res")
        replay("cand")
        first = True
        while True:
            if not first:
                if not len(cand):
                    break
                cp, out_of = cand.pop()
                lb = res.a[0][0]
                if cp > lb:
                    break
                exps += 1
                rest = []
                for s in nbrs[out_of]:
                    if s not in visited:
                        visited.add(s)
                        evals += 1
                        rest.append(s)
            first = False
            for s in rest:
                d = dist[s]
                if out_of is None or d < lb or len(res) < ef:
                    cand.push(d, s)
                    res.push(d, s)
                    if len(res) > ef:
                        res.pop()
        while len(res) > k:
            res.pop()
        out = [None] * len(res)
        for i in range(len(res) - 1, -1, -1):
            out[i] = res.pop()
        return out, evals, exps

    visited = {ep}
    evals, exps = 1, 0
    if not insert(dist[ep], ep):
        return finish_literal(None, 0.0, [ep])          # ST_ENTRY2 with a NaN entrypoint distance
    log.append((dist[ep], ep))
    while True:
        # ---- pick
        idx = next((i for i, e in enumerate(ws) if not e[2]), None)
        found = idx is not None
        cp, cs = (ws[idx][0], ws[idx][1]) if found else (None, None)
        if haz_set and ws and ws[-1][0] < haz_x:
            haz_set = False
        several = found and sum(1 for e in ws if not e[2] and e[0] == cp) > 1
        consult = (has_tie and (several or (haz_set and cp == haz_x))) if found else haz_set
        if consult:
            stats["H1"] += 1
            replay("cand")
            if len(cand):
                cp, cs = cand.pop()
                pos["cand"] = len(log) + 1
                found = not (cp > ws[-1][0])
            else:
                found = False
        if not found:
            break
        for e in ws:
            if e[1] == cs:
                e[2] = True
        lb = ws[-1][0]
        log.append(None)
        exps += 1
        # ---- expand + fold (the kernel test-and-sets the visited bits of the whole list, then folds in list order)
        fresh = [s for s in nbrs[cs] if s not in visited]
        visited.update(fresh)
        evals += len(fresh)
        for i, s in enumerate(fresh):
            d = dist[s]
            if d < lb or len(ws) < ef:
                if d != d:
                    return finish_literal(cs, lb, fresh[i:])
                if not insert(d, s):
                    stats["H2"] += 1
                    replay("res")
                    res.push(d, s)
                    _, ts = res.pop()
                    pos["res"] = len(log) + 1
                    haz_x = ws[-1][0]                      # priority of the evicted member == the largest one left
                    haz_set = True
                    if ts != s:
                        ws[:] = [e for e in ws if e[1] != ts]
                        insert(d, s, check=False)
                log.append((d, s))
    lim = min(k + 1, len(ws))
    if has_tie and any(ws[i][0] == ws[i + 1][0] for i in range(lim - 1)):
        stats["H3"] += 1
        replay("res")
        while len(res) > k:
            res.pop()
        out = [None] * len(res)
        for i in range(len(res) - 1, -1, -1):
            out[i] = res.pop()
        return out, evals, exps
    return [(e[0], e[1]) for e in ws[:k]], evals, exps


def _canon(r):
    return [("nan" if p != p else p, s) for p, s in r[0]], r[1], r[2]


def _graph(rng, n, deg, levels, nan_frac=0.0):
    nbrs = []
    for v in range(n):
        m = int(rng.integers(1, deg + 1))
        c = set(int(x) for x in rng.integers(0, n, size=m)) - {v}
        nbrs.append(sorted(c))                    # ascending neighbour id, the deterministic iteration order
    dist = rng.integers(0, levels, size=n).astype(np.float32)
    dist[rng.random(n) < nan_frac] = np.nan                  # zero-norm rows: 0/0 in the reference too
    return nbrs, [float(x) for x in dist]


@pytest.mark.parametrize("levels", [3, 8, 40, 100000])
def test_register_result_set_with_lazy_literal_heaps_walks_like_go(levels):
    rng = np.random.default_rng(levels)
    stats = {"H1": 0, "H2": 0, "H3": 0}
    for trial in range(1500):
        n = int(rng.integers(2, 120))
        nbrs, dist = _graph(rng, n, int(rng.integers(1, 12)), levels, nan_frac=0.03 if trial % 3 == 0 else 0.0)
        ep = int(rng.integers(0, n))
        ef = int(rng.integers(1, 24))
        k = int(rng.integers(1, ef + 1))
        want = literal_search(nbrs, dist, ep, ef, k)
        got = model_search(nbrs, dist, ep, ef, k, stats)
        assert _canon(got) == _canon(want), (levels, trial, n, ef, k)
    if levels <= 40:
        assert stats["H1"] and stats["H2"] and stats["H3"], stats     # every hazard path was exercised
