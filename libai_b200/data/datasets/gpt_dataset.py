"""GPT-style causal LM dataset over a flattened token stream.

Spec: reference libai/data/datasets/gpt_dataset.py:33-301 (Megatron GPT2Dataset): three cached
index arrays

* ``doc_idx``     – epochs × documents, shuffled (the last epoch is shuffled separately when it is
                    used for < 80 % of its samples),
* ``sample_idx``  – ``[n + 1, 2]`` (position in ``doc_idx``, token offset) from the native helper,
* ``shuffle_idx`` – sample permutation (two independent ranges when the last epoch is separate),

all derived from one ``np.random.RandomState(seed)`` in this order (data-order contract) and cached
as ``<prefix>_<name>_indexmap_<ns>ns_<sl>sl_<seed>s_{doc,sample,shuffle}_idx.npy``.  A sample is
``seq_length + 1`` consecutive tokens: inputs ``[:-1]``, labels ``[1:]`` (labels go to the last
pipeline stage).
"""
from __future__ import annotations

import logging
import os
import time

import numpy as np
import torch

from libai_b200.data.data_utils.dataset_utils import is_shared_folder
from libai_b200.data.structures import DistTensorData, Instance
from libai_b200.utils import distributed as dutil

logger = logging.getLogger(__name__)


class GPT2Dataset(torch.utils.data.Dataset):
    def __init__(self, name, tokenizer, data_prefix, indexed_dataset, max_num_samples, max_seq_length, seed=1234):
        self.name = name
        self.tokenizer = tokenizer
        self.indexed_dataset = indexed_dataset
        documents = np.arange(0, indexed_dataset.sizes.shape[0], dtype=np.int32)
        self.doc_idx, self.sample_idx, self.shuffle_idx = _build_index_mappings(
            name, data_prefix, documents, indexed_dataset.sizes, max_num_samples, max_seq_length, seed
        )

    def __len__(self):
        return self.sample_idx.shape[0] - 1

    def _tokens(self, idx):
        (d0, o0), (d1, o1) = self.sample_idx[idx], self.sample_idx[idx + 1]
        ds = self.indexed_dataset
        if d0 == d1:
            return ds.get(self.doc_idx[d0], offset=int(o0), length=int(o1 - o0 + 1))
        parts = [ds.get(self.doc_idx[d0], offset=int(o0))]
        parts += [ds.get(self.doc_idx[i]) for i in range(d0 + 1, d1)]
        parts.append(ds.get(self.doc_idx[d1], length=int(o1 + 1)))
        return np.concatenate(parts)

    def __getitem__(self, idx):
        sample = np.asarray(self._tokens(int(self.shuffle_idx[idx])), dtype=np.int64)
        return Instance(
            input_ids=DistTensorData(torch.from_numpy(sample[:-1].copy())),
            labels=DistTensorData(torch.from_numpy(sample[1:].copy()), placement_idx=-1),
        )


def _num_tokens(documents, sizes):
    return int(np.sum(sizes[documents]))


def _num_epochs(tokens_per_epoch, seq_length, num_samples):
    """Smallest number of epochs yielding at least ``num_samples`` samples."""
    epochs, tokens = 0, 0
    while True:
        epochs += 1
        tokens += tokens_per_epoch
        if (tokens - 1) // seq_length >= num_samples:
            return epochs


def _build_doc_idx(documents, num_epochs, np_rng, separate_last_epoch):
    if not separate_last_epoch or num_epochs == 1:
        doc_idx = np.tile(np.asarray(documents, dtype=np.int32), num_epochs)
        np_rng.shuffle(doc_idx)
        return doc_idx
    head = _build_doc_idx(documents, num_epochs - 1, np_rng, False)
    tail = _build_doc_idx(documents, 1, np_rng, False)
    return np.concatenate((head, tail))


def _build_shuffle_idx(num_samples, total_size, np_rng):
    logger.info(f" > building shuffle index with split [0, {num_samples}) and [{num_samples}, {total_size}) ...")
    dtype_ = np.int64 if total_size >= (np.iinfo(np.uint32).max - 1) else np.uint32
    first = np.arange(0, num_samples, dtype=dtype_)
    np_rng.shuffle(first)
    if num_samples == total_size:
        return first
    last = np.arange(num_samples, total_size, dtype=dtype_)
    np_rng.shuffle(last)
    return np.concatenate((first, last))


def _build_index_mappings(name, data_prefix, documents, sizes, num_samples, seq_length, seed):
    tokens_per_epoch = _num_tokens(documents, sizes)
    num_epochs = _num_epochs(tokens_per_epoch, seq_length, num_samples)
    np_rng = np.random.RandomState(seed=seed)
    stem = f"{data_prefix}_{name}_indexmap_{num_samples}ns_{seq_length}sl_{seed}s"
    files = {k: f"{stem}_{k}_idx.npy" for k in ("doc", "sample", "shuffle")}
    folder = os.path.dirname(stem) or "."
    builder_rank = dutil.get_rank() if is_shared_folder(folder) else dutil.get_local_rank()

    if builder_rank == 0 and not all(os.path.isfile(f) for f in files.values()):
        logger.info(" > WARNING: could not find index map files, building the indices on rank 0 ...")
        separate_last_epoch = False
        samples_before_last = None
        if num_epochs == 1:
            logger.info(" > only one epoch required, setting separate_last_epoch to False")
        else:
            samples_before_last = ((num_epochs - 1) * tokens_per_epoch - 1) // seq_length
            last_epoch_samples = num_samples - samples_before_last
            per_epoch = (tokens_per_epoch - 1) // seq_length
            assert last_epoch_samples >= 0, "last epoch number of samples should be non-negative."
            assert last_epoch_samples < per_epoch + 1, "last epoch number of samples exceeded max value."
            separate_last_epoch = last_epoch_samples < int(0.80 * per_epoch)
            logger.info(
                f" > last epoch number of samples ({last_epoch_samples}) vs 80% of samples per epoch ({per_epoch}): "
                f"separate_last_epoch={separate_last_epoch}"
            )
        t0 = time.time()
        doc_idx = _build_doc_idx(documents, num_epochs, np_rng, separate_last_epoch)
        np.save(files["doc"], doc_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save doc-idx mapping (seconds): {time.time() - t0:4f}")
        t0 = time.time()
        from libai_b200.data.data_utils import helpers

        assert doc_idx.dtype == np.int32 and sizes.dtype == np.int32
        sample_idx = helpers.build_sample_idx(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch)
        np.save(files["sample"], sample_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save sample-idx mapping (seconds): {time.time() - t0:4f}")
        t0 = time.time()
        n_first = samples_before_last if separate_last_epoch else sample_idx.shape[0] - 1
        shuffle_idx = _build_shuffle_idx(n_first, sample_idx.shape[0] - 1, np_rng)
        np.save(files["shuffle"], shuffle_idx, allow_pickle=True)
        logger.info(f" > elapsed time to build and save shuffle-idx mapping (seconds): {time.time() - t0:4f}")
    dutil.synchronize()

    t0 = time.time()
    doc_idx = np.load(files["doc"], allow_pickle=True, mmap_mode="r")
    sample_idx = np.load(files["sample"], allow_pickle=True, mmap_mode="r")
    shuffle_idx = np.load(files["shuffle"], allow_pickle=True, mmap_mode="r")
    logger.info(f"    loaded indexed file in {time.time() - t0:3.3f} seconds")
    logger.info(f"    total number of samples: {sample_idx.shape[0]}")
    logger.info(f"    total number of epochs: {num_epochs}")
    return doc_idx, sample_idx, shuffle_idx
