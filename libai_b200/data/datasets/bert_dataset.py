"""BERT pre-training dataset (masked LM + sentence-order / next-sentence head).

Spec: reference libai/data/datasets/bert_dataset.py:26-326.  A sample = sentences
``[start, end)`` from the samples mapping; per-sample ``RandomState((seed + idx) % 2**32)`` drives, in
this order: the A/B split point, the 50 % swap (``ns_labels``), truncation (drop front/back), then
the masking routine.  Fields: ``input_ids, attention_mask, tokentype_ids`` (first stage) and
``ns_labels, lm_labels (−1 = not masked), loss_mask`` (last stage).
"""
from __future__ import annotations

import numpy as np
import torch

from libai_b200.data.data_utils.dataset_utils import create_masked_lm_predictions, get_samples_mapping
from libai_b200.data.structures import DistTensorData, Instance


class BertDataset(torch.utils.data.Dataset):
    def __init__(self, name, tokenizer, indexed_dataset, data_prefix, max_num_samples, mask_lm_prob, max_seq_length,
                 short_seq_prob=0.0, seed=1234, binary_head=True, masking_style="bert"):
        self.name, self.seed = name, seed
        self.masked_lm_prob = mask_lm_prob
        self.max_seq_length = max_seq_length
        self.binary_head = binary_head
        self.masking_style = masking_style
        self.indexed_dataset = indexed_dataset
        # 3 positions are reserved for [CLS] a [SEP] b [SEP]
        self.samples_mapping = get_samples_mapping(
            indexed_dataset, data_prefix, None, max_num_samples, max_seq_length - 3, short_seq_prob, seed, name,
            binary_head,
        )
        self.tokenizer = tokenizer
        vocab = tokenizer.get_vocab()
        self.vocab_id_list = list(vocab.values())
        self.vocab_id_to_token_dict = {v: k for k, v in vocab.items()}
        self.cls_id, self.sep_id = tokenizer.cls_token_id, tokenizer.sep_token_id
        self.mask_id, self.pad_id = tokenizer.mask_token_id, tokenizer.pad_token_id

    def __len__(self):
        return self.samples_mapping.shape[0]

    def __getitem__(self, idx):
        start, end, target_len = (int(x) for x in self.samples_mapping[idx])
        sentences = [self.indexed_dataset[i] for i in range(start, end)]
        rng = np.random.RandomState(seed=(self.seed + idx) % 2 ** 32)
        return build_training_sample(
            self.tokenizer, sentences, target_len, self.max_seq_length, self.vocab_id_list,
            self.vocab_id_to_token_dict, self.cls_id, self.sep_id, self.mask_id, self.pad_id, self.masked_lm_prob, rng,
            self.binary_head, masking_style=self.masking_style,
        )


def get_a_and_b_segments(sample, np_rng):
    """Split the sentences into segment A / B at a random boundary; swap them with p = 0.5."""
    n = len(sample)
    assert n > 1, "make sure each sample has at least two sentences."
    cut = np_rng.randint(1, n) if n >= 3 else 1
    a = [t for s in sample[:cut] for t in s]
    b = [t for s in sample[cut:] for t in s]
    swapped = bool(np_rng.random() < 0.5)
    return (b, a, True) if swapped else (a, b, False)


def truncate_segments(tokens_a, tokens_b, len_a, len_b, max_num_tokens, np_rng):
    """Trim the longer segment, token by token from a random end, until the pair fits (in place)."""
    assert len_a > 0
    if len_a + len_b <= max_num_tokens:
        return False
    while len_a + len_b > max_num_tokens:
        if len_a > len_b:
            len_a -= 1
            victim = tokens_a
        else:
            len_b -= 1
            victim = tokens_b
        if np_rng.random() < 0.5:
            del victim[0]
        else:
            victim.pop()
    return True


def create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id):
    """``[CLS] A [SEP] B [SEP]`` with token types 0 / 1."""
    tokens = [cls_id] + list(tokens_a) + [sep_id]
    types = [0] * len(tokens)
    if tokens_b:
        tokens += list(tokens_b) + [sep_id]
        types += [1] * (len(tokens_b) + 1)
    return tokens, types


def pad_and_convert_to_numpy(tokens, tokentypes, masked_positions, masked_labels, pad_id, max_seq_length):
    n = len(tokens)
    pad = max_seq_length - n
    assert pad >= 0 and len(tokentypes) == n and len(masked_positions) == len(masked_labels)
    tokens_np = np.array(tokens + [pad_id] * pad, dtype=np.int64)
    types_np = np.array(tokentypes + [pad_id] * pad, dtype=np.int64)
    padding_mask = np.array([1] * n + [0] * pad, dtype=bool)
    labels = np.full(max_seq_length, -1, dtype=np.int64)
    loss_mask = np.zeros(max_seq_length, dtype=bool)
    for pos, lab in zip(masked_positions, masked_labels):
        assert pos < n
        labels[pos] = lab
        loss_mask[pos] = True
    return tokens_np, types_np, labels, padding_mask, loss_mask


def build_training_sample(tokenizer, sample, target_seq_length, max_seq_length, vocab_id_list, vocab_id_to_token_dict,
                          cls_id, sep_id, mask_id, pad_id, masked_lm_prob, np_rng, binary_head, masking_style="bert"):
    if binary_head:
        assert len(sample) > 1
    assert target_seq_length <= max_seq_length
    if binary_head:
        tokens_a, tokens_b, is_next_random = get_a_and_b_segments(sample, np_rng)
    else:
        tokens_a, tokens_b, is_next_random = [t for s in sample for t in s], [], False
    truncate_segments(tokens_a, tokens_b, len(tokens_a), len(tokens_b), target_seq_length, np_rng)
    tokens, tokentypes = create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id)
    tokens, positions, labels, _, _ = create_masked_lm_predictions(
        tokenizer, tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
        masked_lm_prob * target_seq_length, np_rng, masking_style=masking_style,
    )
    tok, typ, lab, pad_mask, loss_mask = pad_and_convert_to_numpy(tokens, tokentypes, positions, labels, pad_id, max_seq_length)
    return Instance(
        input_ids=DistTensorData(torch.from_numpy(tok)),
        attention_mask=DistTensorData(torch.from_numpy(pad_mask)),
        tokentype_ids=DistTensorData(torch.from_numpy(typ)),
        ns_labels=DistTensorData(torch.tensor(int(is_next_random), dtype=torch.long), placement_idx=-1),
        lm_labels=DistTensorData(torch.from_numpy(lab), placement_idx=-1),
        loss_mask=DistTensorData(torch.from_numpy(loss_mask), placement_idx=-1),
    )
