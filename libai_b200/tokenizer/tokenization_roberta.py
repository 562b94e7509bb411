"""RoBERTa tokenizer (byte-level BPE with BERT-style special tokens).

Spec: reference libai/tokenizer/tokenization_roberta.py:97-328 — ``<s> a </s>`` / ``<s> a </s> b </s>`` when
``add_bos_token`` (:245-271), all-zero token-type ids (:306-328).
"""
from __future__ import annotations

from typing import List, Optional

from .tokenization_gpt2 import _ByteBPETokenizer, bytes_to_unicode, get_pairs  # noqa: F401  (re-exported)

VOCAB_FILES_NAMES = {"vocab_file": "vocab.json", "merges_file": "merges.txt"}
_NAMES = ["roberta-base", "roberta-large", "roberta-large-mnli", "distilroberta-base"]
PRETRAINED_VOCAB_FILES_MAP = {
    "vocab_file": {n: f"https://huggingface.co/{n}/resolve/main/vocab.json" for n in _NAMES},
    "merges_file": {n: f"https://huggingface.co/{n}/resolve/main/merges.txt" for n in _NAMES},
}
PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES = {n: 512 for n in _NAMES}


class RobertaTokenizer(_ByteBPETokenizer):
    vocab_files_names = VOCAB_FILES_NAMES
    pretrained_vocab_files_map = PRETRAINED_VOCAB_FILES_MAP
    max_model_input_sizes = PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES

    def __init__(self, vocab_file, merges_file, errors="replace", bos_token="<s>", eos_token="</s>", sep_token="</s>",
                 cls_token="<s>", unk_token="<unk>", pad_token="<pad>", mask_token="<mask>", add_bos_token=False,
                 **kwargs):
        super().__init__(bos_token=bos_token, eos_token=eos_token, sep_token=sep_token, cls_token=cls_token,
                         unk_token=unk_token, pad_token=pad_token, mask_token=mask_token, **kwargs)
        self._init_bpe(vocab_file, merges_file, errors)
        self.add_bos_token = add_bos_token

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        cls, sep = ([self.cls_token_id], [self.sep_token_id]) if self.add_bos_token else ([], [])
        if token_ids_1 is None:
            return cls + token_ids_0 + sep
        return cls + token_ids_0 + sep + token_ids_1 + sep

    def create_token_type_ids_from_sequences(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        """RoBERTa does not use token types: a list of zeros of the full (``<s> a </s></s> b </s>``) length."""
        if token_ids_1 is None:
            return [0] * (len(token_ids_0) + 2)
        return [0] * (len(token_ids_0) + len(token_ids_1) + 4)
