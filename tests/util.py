"""Shared helpers for the parity tests (seeded inputs per SURVEY §8d: base 0xC0177, queries 0xC0178)."""
import numpy as np

BASE_SEED, QUERY_SEED = 0xC0177, 0xC0178


def rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def uniform(n, d, seed=BASE_SEED):
    """Reference-style data: rand.Float32() uniform [0,1) (compresshelper_test.go:31-37)."""
    return rng(seed).random((n, d), dtype=np.float32)


def normal(n, d, seed=BASE_SEED):
    return rng(seed).standard_normal((n, d)).astype(np.float32)


def sparse_ids(n, seed=BASE_SEED):
    """Snowflake-like sparse unique u64 ids (edge/id_generator.go:24-30)."""
    r = rng(seed ^ 0x5EED)
    ids = np.unique(r.integers(1 << 40, 1 << 62, size=int(n * 1.2) + 8, dtype=np.uint64))
    r.shuffle(ids)
    return ids[:n].copy()


def score_bits(sc):
    """fp32 bit patterns with every NaN canonicalised (NaN payload/sign is not part of the contract:
    the reference's own 0/0 comes out of an x86 divss, ours out of the GPU's divider)."""
    a = np.asarray(sc, np.float32).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a.tobytes()


def assert_same_hits(got_ids, got_sc, want_ids, want_sc, what=""):
    assert len(got_ids) == len(want_ids), f"{what}: count {len(got_ids)} != {len(want_ids)}"
    assert score_bits(got_sc) == score_bits(want_sc), \
        f"{what}: scores differ\n got {got_sc}\nwant {want_sc}"
    assert np.array_equal(np.asarray(got_ids, np.uint64), np.asarray(want_ids, np.uint64)), \
        f"{what}: ids differ\n got {got_ids}\nwant {want_ids}"
