"""RoBERTa pre-training dataset: single-segment masked LM (no sentence-pair head).

Spec: reference libai/data/datasets/roberta_dataset.py:27-219 — samples mapping built with
``binary_head=False`` (one-sentence documents allowed) and ``max_seq_length - 2`` for
``<s> … </s>``; per-sample ``RandomState((seed + idx) % 2**32)``: truncation from a random end, then
masking; fields ``input_ids, attention_mask, tokentype_ids`` / ``lm_labels, loss_mask``.
"""
import numpy as np
import torch

from libai_b200.data.data_utils.dataset_utils import create_masked_lm_predictions, get_samples_mapping
from libai_b200.data.structures import DistTensorData, Instance

from .bert_dataset import pad_and_convert_to_numpy


class RobertaDataset(torch.utils.data.Dataset):
    def __init__(self, name, tokenizer, indexed_dataset, data_prefix, max_num_samples, mask_lm_prob, max_seq_length,
                 short_seq_prob=0.0, seed=1234, masking_style="bert"):
        super().__init__()
        self.name, self.seed = name, seed
        self.masked_lm_prob, self.max_seq_length, self.masking_style = mask_lm_prob, max_seq_length, masking_style
        self.indexed_dataset = indexed_dataset
        self.samples_mapping = get_samples_mapping(
            indexed_dataset, data_prefix, None, max_num_samples, max_seq_length - 2, short_seq_prob, seed, name,
            binary_head=False,
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
            self.tokenizer, sentences, target_len, self.max_seq_length, self.vocab_id_list, self.vocab_id_to_token_dict,
            self.cls_id, self.sep_id, self.mask_id, self.pad_id, self.masked_lm_prob, rng, masking_style=self.masking_style,
        )


def truncate_segments(tokens, len_tokens, max_num_tokens, np_rng):
    """Drop tokens from a random end until the sequence fits (in place)."""
    assert len_tokens > 0
    if len_tokens <= max_num_tokens:
        return False
    for _ in range(len_tokens - max_num_tokens):
        if np_rng.random() < 0.5:
            del tokens[0]
        else:
            tokens.pop()
    return True


def create_tokens_and_tokentypes(tokens, cls_id, sep_id):
    tokens.insert(0, cls_id)
    tokens.append(sep_id)
    return tokens, [0] * len(tokens)


def build_training_sample(tokenizer, sample, target_seq_length, max_seq_length, vocab_id_list, vocab_id_to_token_dict,
                          cls_id, sep_id, mask_id, pad_id, masked_lm_prob, np_rng, masking_style="bert"):
    assert target_seq_length <= max_seq_length
    tokens = [int(t) for s in sample for t in s]
    truncate_segments(tokens, len(tokens), target_seq_length, np_rng)
    tokens, tokentypes = create_tokens_and_tokentypes(tokens, cls_id, sep_id)
    tokens, positions, labels, _, _ = create_masked_lm_predictions(
        tokenizer, tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
        masked_lm_prob * target_seq_length, np_rng, masking_style=masking_style,
    )
    tok, typ, lab, pad_mask, loss_mask = pad_and_convert_to_numpy(tokens, tokentypes, positions, labels, pad_id, max_seq_length)
    return Instance(
        input_ids=DistTensorData(torch.from_numpy(tok)),
        attention_mask=DistTensorData(torch.from_numpy(pad_mask)),
        tokentype_ids=DistTensorData(torch.from_numpy(typ)),
        lm_labels=DistTensorData(torch.from_numpy(lab), placement_idx=-1),
        loss_mask=DistTensorData(torch.from_numpy(loss_mask), placement_idx=-1),
    )
