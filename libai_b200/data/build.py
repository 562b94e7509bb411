"""Dataloader factories.

Spec: reference libai/data/build.py:28-401 — five factories that receive *lazy* dataset / sampler
records (the trainer instantiates ``cfg.dataloader.*`` with ``_recursive_=False``), inject
``micro_batch_size, consumed_samples, data_parallel_rank/size, seed`` into the sampler, build a
``DataLoader`` over a batch sampler with persistent workers, and (image train loader) attach
``mixup_func``.  ``build_nlp_train_val_test_loader`` carves train/valid/test out of one indexed
dataset by slicing its document index according to ``splits``.
"""
from __future__ import annotations

from torch.utils.data import ConcatDataset, DataLoader

from libai_b200.config import LazyCall, instantiate
from libai_b200.config.dictconfig import OmegaConf
from libai_b200.utils import distributed as dutil

from .samplers import CyclicSampler, SingleRoundSampler
from .structures import Instance


def trivial_batch_collator(batch):
    assert isinstance(batch[0], Instance), "batch[0] must be `instance` for trivial batch collator"
    return Instance.stack(batch)


def _as_dataset(dataset, mixer):
    """Instantiate (if lazy) and merge a dataset or list of datasets."""
    dataset = instantiate(dataset)
    if OmegaConf.is_list(dataset):
        dataset = list(dataset)
    elif not isinstance(dataset, list):
        dataset = [dataset]
    return mixer(dataset) if len(dataset) > 1 else dataset[0]


def _make_loader(dataset, sampler_cfg, batch_size, num_workers, seed, collate_fn, consumed_samples=None, **kwargs):
    sampler_cfg.dataset = dataset
    sampler_cfg.micro_batch_size = batch_size
    if consumed_samples is not None:
        sampler_cfg.consumed_samples = consumed_samples
    sampler_cfg.data_parallel_rank = dutil.get_data_parallel_rank()
    sampler_cfg.data_parallel_size = dutil.get_data_parallel_size()
    sampler_cfg.seed = seed
    sampler = instantiate(sampler_cfg)
    kwargs.setdefault("pin_memory", dutil.get_dist_util().device_type == "cuda")
    return DataLoader(
        dataset,
        batch_sampler=sampler,
        num_workers=num_workers,
        persistent_workers=num_workers > 0,
        collate_fn=collate_fn or trivial_batch_collator,
        **kwargs,
    )


def build_nlp_train_val_test_loader(
    dataset,
    splits,
    weights,
    train_val_test_num_samples,
    train_batch_size,
    test_batch_size,
    train_sampler=LazyCall(CyclicSampler)(shuffle=True),
    test_sampler=LazyCall(SingleRoundSampler)(shuffle=False, drop_last=False),
    num_workers=4,
    consumed_samples=0,
    seed=0,
    collate_fn=None,
    dataset_mixer=ConcatDataset,
):
    """Returns ``(train_loader, valid_loader, test_loader)`` built from datasets that only ship a
    single corpus (``splits`` e.g. ``[[949, 50, 1]]``, one entry per dataset)."""
    from .data_utils import get_train_valid_test_split_

    if OmegaConf.is_list(dataset):
        dataset = list(dataset)
    elif not isinstance(dataset, list):
        dataset = [dataset]
    assert len(dataset) == len(splits), "datasets length must equal splits length"
    assert len(dataset) == len(weights), "datasets length must equal weights length"

    buckets = ([], [], [])
    for lazy_ds, split in zip(dataset, splits):
        indexed = instantiate(lazy_ds.indexed_dataset)
        full_doc_idx = indexed.get_doc_idx()
        n_docs = full_doc_idx.shape[0] - 1
        bounds = get_train_valid_test_split_(n_docs, split)
        for which in range(3):
            indexed.set_doc_idx(full_doc_idx[bounds[which] : bounds[which + 1] + 1])
            lazy_ds.indexed_dataset = indexed
            lazy_ds.max_num_samples = train_val_test_num_samples[which]
            buckets[which].append(instantiate(lazy_ds))
            indexed.set_doc_idx(full_doc_idx)
        assert indexed.doc_idx[0] == 0 and indexed.doc_idx.shape[0] == n_docs + 1
    train_ds, val_ds, test_ds = (dataset_mixer(b) for b in buckets)

    train_loader, _, _ = build_nlp_train_loader(
        dataset=train_ds, train_batch_size=train_batch_size, test_batch_size=None, sampler=train_sampler,
        num_workers=num_workers, consumed_samples=consumed_samples, seed=seed, collate_fn=collate_fn,
    )
    make_eval = lambda ds: build_nlp_test_loader(  # noqa: E731
        dataset=ds, test_batch_size=test_batch_size, sampler=test_sampler.copy() if hasattr(test_sampler, "copy") else test_sampler,
        num_workers=num_workers, seed=seed, collate_fn=collate_fn,
    )
    return train_loader, make_eval(val_ds), make_eval(test_ds)


def build_nlp_train_loader(
    dataset,
    train_batch_size,
    test_batch_size=None,
    sampler=LazyCall(CyclicSampler)(shuffle=True),
    num_workers=4,
    consumed_samples=0,
    seed=0,
    collate_fn=None,
    dataset_mixer=ConcatDataset,
    **kwargs,
):
    """Returns ``(train_loader, None, None)``."""
    ds = _as_dataset(dataset, dataset_mixer)
    loader = _make_loader(ds, sampler, train_batch_size, num_workers, seed, collate_fn, consumed_samples, **kwargs)
    return loader, None, None


def build_nlp_test_loader(
    dataset,
    test_batch_size,
    sampler=LazyCall(SingleRoundSampler)(shuffle=False, drop_last=False),
    num_workers=4,
    seed=0,
    collate_fn=None,
):
    ds = instantiate(dataset)
    return _make_loader(ds, sampler, test_batch_size, num_workers, seed, collate_fn)


def build_image_train_loader(
    dataset,
    train_batch_size,
    test_batch_size=None,
    sampler=LazyCall(CyclicSampler)(shuffle=True),
    num_workers=4,
    consumed_samples=0,
    seed=0,
    collate_fn=None,
    dataset_mixer=ConcatDataset,
    mixup_func=None,
    **kwargs,
):
    """Returns ``(train_loader, None, None)``; ``train_loader.mixup_func`` holds the (instantiated)
    Mixup/CutMix callable applied on device in ``DefaultTrainer.get_batch``."""
    ds = _as_dataset(dataset, dataset_mixer)
    loader = _make_loader(ds, sampler, train_batch_size, num_workers, seed, collate_fn, consumed_samples, **kwargs)
    loader.mixup_func = instantiate(mixup_func)
    return loader, None, None


def build_image_test_loader(
    dataset,
    test_batch_size,
    sampler=LazyCall(SingleRoundSampler)(shuffle=True, drop_last=False),
    num_workers=4,
    seed=0,
    collate_fn=None,
    **kwargs,
):
    ds = instantiate(dataset)
    return _make_loader(ds, sampler, test_batch_size, num_workers, seed, collate_fn, **kwargs)
