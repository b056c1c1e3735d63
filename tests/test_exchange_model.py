"""CPU model of the peer-memory exchange of csrc/comm.cu (the flag hand-shake that replaces ncclAllGather): every rank runs
the same stream-ordered program per step s = 1, 2, ...

    write   own buffer[s & 1]            <- the local search's hits (exchange_begin: slot = seq & 1)
    signal  own flag[s & 1] = s          <- xchg_signal_kernel (release)
    wait    peers' flag[s & 1] >= s      <- xchg_wait_kernel (acquire; compared as (int32)(v - s) >= 0)
    merge   read every rank's buffer[s & 1]   <- K5 through list_bases[]

The claim under test is the one the double-buffering rests on: a rank can only overwrite slot (s & 1) — at step s + 2 —
after every peer has finished merging step s out of it.  The model executes the ranks' programs under arbitrary
interleavings (exhaustively for small cases, randomly for larger ones) and fails if a merge ever observes anything but
step s's data, or if the system deadlocks.  A one-buffer variant of the same protocol is shown to be unsafe, so the test
can tell the difference.  Host logic only; no GPU."""
import itertools
import random

import pytest

WRITE, SIGNAL, WAIT, MERGE = range(4)


def program(steps):
    return [(op, s) for s in range(1, steps + 1) for op in (WRITE, SIGNAL, WAIT, MERGE)]


class World:
    def __init__(self, n_ranks, steps, n_slots=2, merge_reads_per_op=1):
        self.n, self.slots = n_ranks, n_slots
        self.pc = [0] * n_ranks
        self.prog = program(steps)
        self.buf = [[0] * n_slots for _ in range(n_ranks)]      # the step whose data a slot holds
        self.flag = [[0] * n_slots for _ in range(n_ranks)]
        # a merge is not atomic: it reads the peers one after the other; model that as one sub-step per peer
        self.merge_pos = [0] * n_ranks

    def done(self, r):
        return self.pc[r] == len(self.prog)

    def runnable(self, r):
        if self.done(r):
            return False
        op, s = self.prog[self.pc[r]]
        if op == WAIT:
            slot = s % self.slots
            return all(((self.flag[p][slot] - s) & 0xFFFFFFFF) < 0x80000000 for p in range(self.n))
        return True

    def step(self, r):
        """Execute the next operation of rank r (the caller checked runnable).  Returns an error string or None."""
        op, s = self.prog[self.pc[r]]
        slot = s % self.slots
        if op == WRITE:
            self.buf[r][slot] = s
        elif op == SIGNAL:
            self.flag[r][slot] = s
        elif op == MERGE:
            p = self.merge_pos[r]
            if self.buf[p][slot] != s:
                return f"rank {r} merging step {s} read rank {p}'s slot {slot} holding step {self.buf[p][slot]}"
            self.merge_pos[r] += 1
            if self.merge_pos[r] < self.n:
                return None                  # the merge continues with the next peer
            self.merge_pos[r] = 0
        self.pc[r] += 1
        return None

    def key(self):
        return (tuple(self.pc), tuple(self.merge_pos), tuple(map(tuple, self.buf)), tuple(map(tuple, self.flag)))


def explore_exhaustively(n_ranks, steps, n_slots):
    """DFS over every interleaving (memoised on the full state).  Returns (states, first error or None)."""
    import copy
    start = World(n_ranks, steps, n_slots)
    seen, stack = set(), [start]
    while stack:
        w = stack.pop()
        k = w.key()
        if k in seen:
            continue
        seen.add(k)
        ready = [r for r in range(n_ranks) if w.runnable(r)]
        if not ready:
            if not all(w.done(r) for r in range(n_ranks)):
                return len(seen), f"deadlock at pcs {w.pc}"
            continue
        for r in ready:
            w2 = copy.deepcopy(w)
            err = w2.step(r)
            if err:
                return len(seen), err
            stack.append(w2)
    return len(seen), None


def run_random(n_ranks, steps, n_slots, seed, bias=None):
    rng = random.Random(seed)
    w = World(n_ranks, steps, n_slots)
    while True:
        ready = [r for r in range(n_ranks) if w.runnable(r)]
        if not ready:
            return None if all(w.done(r) for r in range(n_ranks)) else f"deadlock at pcs {w.pc}"
        if bias is not None and bias in ready and rng.random() < 0.9:
            r = bias                          # one rank racing ahead as far as the protocol lets it
        else:
            r = rng.choice(ready)
        err = w.step(r)
        if err:
            return err


@pytest.mark.parametrize("n_ranks,steps", [(2, 4), (3, 3)])
def test_double_buffered_exchange_is_safe_under_every_interleaving(n_ranks, steps):
    states, err = explore_exhaustively(n_ranks, steps, n_slots=2)
    assert err is None, err
    assert states > 100


def test_double_buffered_exchange_random_schedules_eight_ranks():
    for seed in range(60):
        assert run_random(8, 12, 2, seed) is None
        assert run_random(8, 12, 2, seed, bias=seed % 8) is None      # a rank that runs ahead whenever it can


def test_a_single_buffer_would_be_overwritten_under_a_reader():
    """The same hand-shake over ONE buffer is broken (a fast rank refills it for step s + 1 while a slow peer is still merging
    step s): the model must find that, otherwise it proves nothing about the two-buffer version."""
    _, err = explore_exhaustively(2, 3, n_slots=1)
    assert err is not None and "holding step" in err


def test_sequence_comparison_survives_wraparound():
    """xchg_wait_kernel compares (int32)(v - seq) >= 0, so the 32-bit step counter may wrap."""
    def arrived(v, seq):
        return ((v - seq) & 0xFFFFFFFF) < 0x80000000
    assert arrived(5, 5) and arrived(6, 5) and not arrived(4, 5)
    assert arrived(1, 0xFFFFFFFF) and arrived(0, 0xFFFFFFFF) is True          # 0 and 1 come after 0xffffffff
    assert not arrived(0xFFFFFFFE, 0xFFFFFFFF)
    for a, b in itertools.product([0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF], repeat=2):
        d = (a - b) & 0xFFFFFFFF
        assert arrived(a, b) == (d < 0x80000000)
