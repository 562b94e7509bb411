"""Unsupervised T5 corpus + on-the-fly span corruption collator.

Spec: reference projects/T5/datasets/dataset.py — ``UnsuperviseT5Dataset`` (:87-101, list of token-id sequences),
``compute_input_and_target_lengths`` (:39-84), ``collate_fn`` (:104-262): random span noise mask with mean span
length, sentinel replacement counting down from ``vocab_size-1``, EOS appended, decoder inputs = targets shifted
right with ``decoder_start_token_id``, plus the three attention masks and the loss mask.
"""
from __future__ import annotations

import json
import os
from typing import List

import numpy as np
import torch
from torch.utils.data import Dataset

from libai_b200.data.structures import DistTensorData, Instance


def get_data(path) -> List[List[int]]:
    """A ``.json``/``.jsonl`` file (or a directory of them) with one list of token ids per line / entry."""
    files = sorted(os.path.join(path, f) for f in os.listdir(path)) if os.path.isdir(path) else [path]
    out: List[List[int]] = []
    for fp in files:
        with open(fp, "r", encoding="utf-8") as f:
            text = f.read().strip()
        if text.startswith("[["):
            out.extend(json.loads(text))
        else:
            out.extend(json.loads(ln) for ln in text.splitlines() if ln.strip())
    return out


def compute_input_and_target_lengths(inputs_length, noise_density, mean_noise_span_length):
    """Raw token count that, after span corruption, yields exactly ``inputs_length`` encoder tokens; returns
    ``(tokens_length, targets_length)``."""

    def lengths(tokens_length):
        num_noise = int(round(tokens_length * noise_density))
        num_spans = int(round(num_noise / mean_noise_span_length))
        # every span becomes one sentinel on each side; both sides get an EOS
        return tokens_length - num_noise + num_spans + 1, num_noise + num_spans + 1

    tokens_length = inputs_length
    while lengths(tokens_length + 1)[0] <= inputs_length:
        tokens_length += 1
    inputs_len, targets_len = lengths(tokens_length)
    if noise_density == 0.5 and targets_len > inputs_len:
        tokens_length -= 1
        targets_len -= 1
    return tokens_length, targets_len


class UnsuperviseT5Dataset(Dataset):
    def __init__(self, data_path):
        self.data = get_data(data_path) if isinstance(data_path, (str, os.PathLike)) else list(data_path)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return {"input_ids": self.data[index]}


class collate_fn:
    def __init__(self, vocab_size, max_seq_length, noise_density, mean_noise_span_length, eos_token_id=1, pad_token_id=0,
                 decoder_start_token_id=0, seed=None):
        self.vocab_size, self.max_seq_length = vocab_size, max_seq_length
        self.noise_density, self.mean_noise_span_length = noise_density, mean_noise_span_length
        self.eos_token_id, self.pad_token_id, self.decoder_start_token_id = eos_token_id, pad_token_id, decoder_start_token_id
        self.expanded_inputs_length, self.target_length = compute_input_and_target_lengths(
            max_seq_length, noise_density, mean_noise_span_length)
        self.rng = np.random.RandomState(seed)

    # ---- span corruption ---------------------------------------------------------------------------
    def random_spans_noise_mask(self, length):
        num_noise = min(max(int(round(length * self.noise_density)), 1), length - 1)
        num_spans = max(int(round(num_noise / self.mean_noise_span_length)), 1)
        num_keep = length - num_noise

        def segmentation(num_items, num_segments):
            marks = np.arange(num_items - 1) < (num_segments - 1)
            self.rng.shuffle(marks)
            seg_id = np.cumsum(np.pad(marks, [[1, 0]]))
            return np.bincount(seg_id, minlength=num_segments)

        noise, keep = segmentation(num_noise, num_spans), segmentation(num_keep, num_spans)
        interleaved = np.reshape(np.stack([keep, noise], axis=1), [num_spans * 2])
        starts = np.cumsum(interleaved)[:-1]
        indicator = np.zeros((length,), dtype=np.int8)
        indicator[starts] = 1
        return (np.cumsum(indicator) % 2).astype(bool)

    def create_sentinel_ids(self, mask):
        start = mask & ~np.roll(mask, 1, axis=-1)
        start[:, 0] = mask[:, 0]
        ids = np.where(start, np.cumsum(start, axis=-1), 0)
        ids = np.where(ids != 0, self.vocab_size - ids, 0)
        return ids - (mask & ~start)  # −1 marks the rest of a span (dropped)

    def filter_input_ids(self, input_ids, sentinel_ids):
        fused = np.where(sentinel_ids != 0, sentinel_ids, input_ids)
        rows = [row[row >= 0] for row in fused]
        out = np.stack(rows)
        return np.concatenate([out, np.full((out.shape[0], 1), self.eos_token_id, dtype=np.int64)], axis=-1)

    def shift_tokens_right(self, labels):
        shifted = np.zeros_like(labels)
        shifted[:, 1:] = labels[:, :-1]
        shifted[:, 0] = self.decoder_start_token_id
        return np.where(shifted == -100, self.pad_token_id, shifted)

    def __call__(self, examples):
        need = self.expanded_inputs_length
        rows = []
        for ex in examples:
            ids = list(ex["input_ids"])[:need]
            rows.append(ids + [self.pad_token_id] * (need - len(ids)))
        input_ids = np.asarray(rows, dtype=np.int64)
        mask = np.stack([self.random_spans_noise_mask(need) for _ in range(len(rows))])
        enc = self.filter_input_ids(input_ids, self.create_sentinel_ids(mask.astype(np.int8).astype(bool)))
        labels = self.filter_input_ids(input_ids, self.create_sentinel_ids(~mask))
        assert enc.shape[-1] == self.max_seq_length and labels.shape[-1] == self.target_length, (enc.shape, labels.shape)
        dec = self.shift_tokens_right(labels)
        b, s, t = enc.shape[0], enc.shape[1], dec.shape[1]
        enc_t, dec_t, lab_t = torch.from_numpy(enc), torch.from_numpy(dec), torch.from_numpy(labels)
        enc_pad = enc_t != self.pad_token_id
        return Instance(
            encoder_input_ids=DistTensorData(enc_t),
            decoder_input_ids=DistTensorData(dec_t),
            encoder_attn_mask=DistTensorData((enc_pad[:, :, None] & enc_pad[:, None, :])),
            decoder_attn_mask=DistTensorData(torch.ones(b, t, t, dtype=torch.bool).tril()),
            encoder_decoder_attn_mask=DistTensorData(enc_pad[:, None, :].expand(b, t, s).contiguous()),
            lm_labels=DistTensorData(lab_t, placement_idx=-1),
            loss_mask=DistTensorData((lab_t != self.pad_token_id).long(), placement_idx=-1),
        )
