"""T5 span-corruption dataset.

Spec: reference libai/data/datasets/t5_dataset.py:28-347 — samples mapping with
``binary_head=False`` and ``max_seq_length - 2``; per-sample ``RandomState(seed + idx)``; masking with
``max_ngrams=10, geometric_dist=True, masking_style="t5"``; every masked span is replaced by one
sentinel token in the encoder input and emitted as ``sentinel + span`` in the decoder stream
(decoder input starts with BOS, labels end with EOS).  Fields: ``encoder_input_ids,
decoder_input_ids, encoder_attn_mask [s,s], decoder_attn_mask [d,d] (causal), encoder_decoder_attn_mask
[d,s]`` / ``lm_labels, loss_mask`` (last stage).
"""
import collections

import numpy as np
import torch

from libai_b200.data.data_utils.dataset_utils import create_masked_lm_predictions, get_samples_mapping
from libai_b200.data.structures import DistTensorData, Instance


class T5Dataset(torch.utils.data.Dataset):
    def __init__(self, name, tokenizer, indexed_dataset, data_prefix, max_num_samples, masked_lm_prob, max_seq_length,
                 max_seq_length_dec, short_seq_prob, seed):
        self.name, self.seed = name, seed
        self.masked_lm_prob = masked_lm_prob
        self.max_seq_length, self.max_seq_length_dec = max_seq_length, max_seq_length_dec
        self.indexed_dataset = indexed_dataset
        self.samples_mapping = get_samples_mapping(
            indexed_dataset, data_prefix, None, max_num_samples, max_seq_length - 2, short_seq_prob, seed, name, False
        )
        self.tokenizer = tokenizer
        tokenizer.add_tokens([tokenizer._bos_token, tokenizer._eos_token, *tokenizer._additional_special_tokens])
        vocab = tokenizer.get_vocab()
        self.vocab_id_to_token_dict = {v: k for k, v in vocab.items()}
        self.vocab_id_list = list(self.vocab_id_to_token_dict.keys())
        self.cls_id, self.sep_id = vocab[tokenizer._cls_token], vocab[tokenizer._sep_token]
        self.mask_id, self.pad_id = vocab[tokenizer._mask_token], vocab[tokenizer._pad_token]
        self.bos_id, self.eos_id = vocab[tokenizer._bos_token], vocab[tokenizer._eos_token]
        self.sentinel_tokens = [vocab[x] for x in tokenizer._additional_special_tokens]
        assert len(self.sentinel_tokens) > 0

    def __len__(self):
        return self.samples_mapping.shape[0]

    def __getitem__(self, idx):
        start, end, target_len = (int(x) for x in self.samples_mapping[idx])
        sentences = [self.indexed_dataset[i] for i in range(start, end)]
        rng = np.random.RandomState(seed=self.seed + idx)
        return build_training_sample(
            self.tokenizer, sentences, target_len, self.max_seq_length, self.max_seq_length_dec, self.vocab_id_list,
            self.vocab_id_to_token_dict, self.cls_id, self.sep_id, self.mask_id, self.pad_id, self.masked_lm_prob, rng,
            self.bos_id, self.eos_id, self.sentinel_tokens,
        )


def make_attention_mask(source_block, target_block):
    """``[len(source), len(target)]``: 1 where both positions hold real tokens (id >= 1)."""
    return ((target_block[None, :] >= 1) & (source_block[:, None] >= 1)).astype(np.int64)


def make_history_mask(block):
    n = block.shape[0]
    pos = np.arange(n)
    return (pos[None, :] <= pos[:, None]).astype(np.int64)


def pad_and_convert_to_numpy(tokens, masked_positions, masked_labels, pad_id, max_seq_length, max_seq_length_dec,
                             masked_spans=None, bos_id=None, eos_id=None, sentinel_tokens=None):
    """Assemble encoder / decoder streams from the masked spans and pad them."""
    sentinels = collections.deque(sentinel_tokens)
    enc, dec_in, dec_out = [], [bos_id], []
    cursor = 0
    for span in masked_spans:
        flag = sentinels.popleft()
        dec_in += [flag, *span.label]
        dec_out += [flag, *span.label]
        enc += tokens[cursor : span.index[0]] + [flag]
        cursor = span.index[-1] + 1
    dec_out.append(eos_id)
    enc += tokens[cursor:]

    pad_enc = max_seq_length - len(enc)
    pad_dec = max_seq_length_dec - len(dec_in)
    assert pad_enc >= 0 and pad_dec >= 0 and len(masked_positions) == len(masked_labels)
    tokens_enc = np.array(enc + [pad_id] * pad_enc, dtype=np.int64)
    tokens_dec = np.array(dec_in + [pad_id] * pad_dec, dtype=np.int64)
    enc_mask = make_attention_mask(tokens_enc, tokens_enc)
    enc_dec_mask = make_attention_mask(tokens_dec, tokens_enc)
    dec_mask = make_attention_mask(tokens_dec, tokens_dec) * make_history_mask(tokens_dec)
    labels = np.array(dec_out + [-1] * pad_dec, dtype=np.int64)
    loss_mask = np.array([1] * len(dec_in) + [0] * pad_dec, dtype=bool)
    t = torch.from_numpy
    return (t(tokens_enc), t(tokens_dec), t(labels), t(enc_mask).bool(), t(dec_mask).bool(), t(enc_dec_mask).bool(),
            t(loss_mask))


def build_training_sample(tokenizer, sample, target_seq_length, max_seq_length, max_seq_length_dec, vocab_id_list,
                          vocab_id_to_token_dict, cls_id, sep_id, mask_id, pad_id, masked_lm_prob, np_rng, bos_id=None,
                          eos_id=None, sentinel_tokens=None):
    assert target_seq_length <= max_seq_length
    tokens = [int(t) for s in sample for t in s][:target_seq_length]
    tokens, positions, labels, _, spans = create_masked_lm_predictions(
        tokenizer, tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
        masked_lm_prob * target_seq_length, np_rng, max_ngrams=10, geometric_dist=True, masking_style="t5",
    )
    enc, dec_in, lab, enc_mask, dec_mask, enc_dec_mask, loss_mask = pad_and_convert_to_numpy(
        tokens, positions, labels, pad_id, max_seq_length, max_seq_length_dec, spans, bos_id, eos_id, sentinel_tokens
    )
    return Instance(
        encoder_input_ids=DistTensorData(enc),
        decoder_input_ids=DistTensorData(dec_in),
        encoder_attn_mask=DistTensorData(enc_mask),
        decoder_attn_mask=DistTensorData(dec_mask),
        encoder_decoder_attn_mask=DistTensorData(enc_dec_mask),
        lm_labels=DistTensorData(lab, placement_idx=-1),
        loss_mask=DistTensorData(loss_mask, placement_idx=-1),
    )
