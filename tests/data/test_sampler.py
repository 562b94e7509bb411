"""Sampler semantics (the reference documents them in a commented-out test: tests/data/test_sampler.py):
contiguous per-rank buckets, seeded per-epoch shuffle, exact resume from ``consumed_samples``, single-round
evaluation with remainder handling."""
import itertools

from libai_b200.data.samplers import CyclicSampler, SingleRoundSampler


def _take(sampler, n):
    return list(itertools.islice(iter(sampler), n))


def test_cyclic_sampler_iterates_and_wraps():
    ds = list(range(10))
    s = CyclicSampler(ds, micro_batch_size=4, shuffle=False)
    batches = _take(s, 5)
    flat = [i for b in batches for i in b]
    # an epoch is the largest multiple of the global batch (10 // 4 * 4 = 8 samples); the remainder is dropped
    assert all(len(b) == 4 for b in batches) and flat[:8] == list(range(8)) and flat[8:12] == [0, 1, 2, 3]


def test_cyclic_sampler_dp_shards_are_disjoint_and_cover():
    ds = list(range(32))
    per_rank = [_take(CyclicSampler(ds, micro_batch_size=2, shuffle=True, data_parallel_rank=r, data_parallel_size=4, seed=3), 4)
                for r in range(4)]
    seen = sorted(i for batches in per_rank for b in batches for i in b)
    assert seen == list(range(32))


def test_cyclic_sampler_resume_is_exact():
    ds = list(range(50))
    full = _take(CyclicSampler(ds, micro_batch_size=3, shuffle=True, data_parallel_rank=1, data_parallel_size=2, seed=7), 12)
    # 5 steps consumed globally = 5 * 3 * 2 samples
    resumed = _take(CyclicSampler(ds, micro_batch_size=3, shuffle=True, consumed_samples=5 * 3 * 2, data_parallel_rank=1,
                                  data_parallel_size=2, seed=7), 7)
    assert resumed == full[5:]


def test_cyclic_sampler_epoch_reshuffle():
    ds = list(range(8))
    s = _take(CyclicSampler(ds, micro_batch_size=8, shuffle=True, seed=1), 2)
    assert sorted(s[0]) == sorted(s[1]) == list(range(8)) and s[0] != s[1]


def test_single_round_sampler():
    ds = list(range(10))
    out = list(SingleRoundSampler(ds, micro_batch_size=4, shuffle=False, drop_last=False))
    assert out == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    out = list(SingleRoundSampler(ds, micro_batch_size=4, shuffle=False, drop_last=True))
    assert out == [[0, 1, 2, 3], [4, 5, 6, 7]]
    parts = [list(SingleRoundSampler(ds, micro_batch_size=2, data_parallel_rank=r, data_parallel_size=2)) for r in range(2)]
    assert sorted(i for p in parts for b in p for i in b) == list(range(10))
