"""Padding / causal masks for the couplet seq2seq batches (reference projects/Couplets/dataset/mask.py)."""
import numpy as np


def make_padding_mask(q_ids, kv_ids, pad_id=0):
    q = (np.array(q_ids) != pad_id).reshape(-1, 1)
    kv = (np.array(kv_ids) != pad_id).reshape(1, -1)
    return (q * kv).astype(np.int64)


def make_sequence_mask(ids):
    n = len(ids)
    return np.tril(np.ones((n, n), dtype=np.int64))
