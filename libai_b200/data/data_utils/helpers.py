"""numpy front-end of the native index builders (same call signatures as the reference's pybind
module ``libai.data.data_utils.helpers``: helpers.cpp:602-607)."""
from __future__ import annotations

import ctypes
import logging

import numpy as np

from . import helpers_build

_LIB = None
logger = logging.getLogger(__name__)


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(helpers_build.ensure_built())
        lib.lb_num_samples.restype = ctypes.c_int64
        lib.lb_num_samples.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64]
        lib.lb_build_sample_idx.restype = None
        lib.lb_build_mapping.restype = ctypes.c_uint64
        lib.lb_build_blocks_mapping.restype = ctypes.c_uint64
        lib.lb_build_blending_indices.restype = None
        _LIB = lib
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def build_sample_idx(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch):
    """int32 ``[num_samples + 1, 2]``: (index into doc_idx, token offset) of every sample start."""
    assert seq_length > 1 and num_epochs > 0 and tokens_per_epoch > 1
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    doc_idx = np.ascontiguousarray(doc_idx, dtype=np.int32)
    lib = _lib()
    n = lib.lb_num_samples(int(seq_length), int(num_epochs), int(tokens_per_epoch))
    out = np.empty((n + 1, 2), dtype=np.int32)
    lib.lb_build_sample_idx(_ptr(sizes), _ptr(doc_idx), ctypes.c_int32(seq_length), ctypes.c_int32(num_epochs),
                            ctypes.c_int64(tokens_per_epoch), _ptr(out))
    return out


def build_mapping(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose=False,
                  min_num_sent=2):
    """``[n, 3]`` rows (first sentence, last sentence + 1, target length), shuffled."""
    assert num_epochs > 0 and max_seq_length > 1 and 0.0 <= short_seq_prob <= 1.0 and seed > 0
    docs = np.ascontiguousarray(docs, dtype=np.int64)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    use_u64 = int(sizes.size > np.iinfo(np.uint32).max)
    dtype = np.uint64 if use_u64 else np.uint32
    lib = _lib()
    args = [_ptr(docs), ctypes.c_int64(docs.shape[0] - 1), _ptr(sizes), ctypes.c_int32(num_epochs),
            ctypes.c_uint64(int(max_num_samples)), ctypes.c_int32(max_seq_length), ctypes.c_double(short_seq_prob),
            ctypes.c_int32(seed), ctypes.c_int32(min_num_sent), ctypes.c_int32(use_u64)]
    n = lib.lb_build_mapping(*args, None)
    out = np.empty((n, 3), dtype=dtype)
    lib.lb_build_mapping(*args, _ptr(out))
    if verbose:
        logger.info(f"    built {n} sentence-span samples (seed {seed}, max_seq_length {max_seq_length})")
    return out


def build_blocks_mapping(docs, sizes, titles_sizes, num_epochs, max_num_samples, max_seq_length, seed, verbose=False,
                         use_one_sent_blocks=False):
    """``[n, 4]`` rows (first sentence, last sentence + 1, document, block id), shuffled."""
    docs = np.ascontiguousarray(docs, dtype=np.int64)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    titles_sizes = np.ascontiguousarray(titles_sizes, dtype=np.int32)
    use_u64 = int(sizes.size > np.iinfo(np.uint32).max)
    dtype = np.uint64 if use_u64 else np.uint32
    lib = _lib()
    args = [_ptr(docs), ctypes.c_int64(docs.shape[0] - 1), _ptr(sizes), _ptr(titles_sizes), ctypes.c_int32(num_epochs),
            ctypes.c_uint64(int(max_num_samples)), ctypes.c_int32(max_seq_length), ctypes.c_int32(seed),
            ctypes.c_int32(int(use_one_sent_blocks)), ctypes.c_int32(use_u64)]
    n = lib.lb_build_blocks_mapping(*args, None)
    out = np.empty((n, 4), dtype=dtype)
    lib.lb_build_blocks_mapping(*args, _ptr(out))
    return out


def build_blending_indices(dataset_index, dataset_sample_index, weights, num_datasets, size, verbose=False):
    """In place: ``dataset_index`` (uint8) / ``dataset_sample_index`` (int64) of length ``size``."""
    assert dataset_index.dtype == np.uint8 and dataset_sample_index.dtype == np.int64 and num_datasets <= 256
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    _lib().lb_build_blending_indices(_ptr(dataset_index), _ptr(dataset_sample_index), _ptr(weights),
                                     ctypes.c_int32(num_datasets), ctypes.c_int64(size))
