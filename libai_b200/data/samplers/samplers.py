"""Batch samplers.

Spec: reference libai/data/samplers/samplers.py — ``CyclicSampler`` (:20-106; infinite, every DP
rank owns the contiguous bucket ``[rank·E, (rank+1)·E)`` with ``E = data_size // (mb·D) · mb``,
per-epoch reshuffle seeded ``seed + epoch``, exact resume through ``consumed_samples``) and
``SingleRoundSampler`` (:109-185; one pass for evaluation, remainder spread over the first ranks,
short ranks padded with index 0 so every rank yields the same number of batches).
The commented-out specs of tests/data/test_sampler.py hold for this implementation.
"""
import torch
from torch.utils.data import Sampler


class CyclicSampler(Sampler):
    def __init__(self, dataset, micro_batch_size, shuffle=False, consumed_samples=0, data_parallel_rank=0,
                 data_parallel_size=1, seed=0):
        self.dataset = dataset
        self.data_size = len(dataset)
        self.shuffle = shuffle
        self.data_parallel_rank = data_parallel_rank
        self.data_parallel_size = data_parallel_size
        self.micro_batch_size = micro_batch_size
        self.actual_batch_size = micro_batch_size * data_parallel_size
        self.data_size_per_epoch = self.data_size // self.actual_batch_size * micro_batch_size
        assert self.data_size_per_epoch > 0, (
            f"dataset of {self.data_size} samples is smaller than one global batch ({self.actual_batch_size})"
        )
        self.consumed_samples = consumed_samples
        self.seed = seed

    def _epoch_order(self, epoch: int):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + epoch)
            return torch.randperm(self.data_size_per_epoch, generator=g).tolist()
        return list(range(self.data_size_per_epoch))

    def __iter__(self):
        per_rank_consumed = self.consumed_samples // self.data_parallel_size
        epoch = per_rank_consumed // self.data_size_per_epoch
        skip = per_rank_consumed % self.data_size_per_epoch
        base = self.data_parallel_rank * self.data_size_per_epoch
        batch = []
        while True:
            order = self._epoch_order(epoch)[skip:]
            indices = [base + i for i in order]
            epoch += 1
            skip = 0
            if getattr(self.dataset, "supports_prefetch", False):
                self.dataset.prefetch(indices)
            for idx in indices:
                batch.append(idx)
                if len(batch) == self.micro_batch_size:
                    self.consumed_samples += self.actual_batch_size
                    yield batch
                    batch = []

    def __len__(self):
        return self.data_size

    def set_consumed_samples(self, consumed_samples):
        """Resume point: number of *global* samples already trained on."""
        self.consumed_samples = consumed_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


class SingleRoundSampler(Sampler):
    def __init__(self, dataset, micro_batch_size, shuffle=False, data_parallel_rank=0, data_parallel_size=1,
                 seed=0, drop_last=False):
        self.dataset = dataset
        self.data_size = len(dataset)
        self.shuffle = shuffle
        self.data_parallel_rank = data_parallel_rank
        self.data_parallel_size = data_parallel_size
        self.micro_batch_size = micro_batch_size
        self.seed = seed
        self.drop_last = drop_last

    def __iter__(self):
        bucket, remain = divmod(self.data_size, self.data_parallel_size)
        start = self.data_parallel_rank * bucket + min(self.data_parallel_rank, remain)
        if self.data_parallel_rank < remain:
            bucket += 1
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed)
            order = torch.randperm(bucket, generator=g).tolist()
        else:
            order = range(bucket)
        indices = [start + i for i in order]
        if getattr(self.dataset, "supports_prefetch", False):
            self.dataset.prefetch(indices)
        batch = []
        for idx in indices:
            batch.append(idx)
            if len(batch) == self.micro_batch_size:
                yield batch
                batch = []
        if not self.drop_last:
            if self.data_parallel_rank >= remain and remain > 0:
                batch.append(0)  # pad so that all ranks iterate the same number of samples
            if batch:
                yield batch

    def __len__(self):
        gbs = self.micro_batch_size * self.data_parallel_size
        return self.data_size // gbs if self.drop_last else (self.data_size + gbs - 1) // gbs
