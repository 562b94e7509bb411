"""T5 SentencePiece tokenizer.

Spec: reference libai/tokenizer/tokenization_t5.py:42-202 — sentencepiece model + ``extra_ids`` sentinel tokens
``<extra_id_N>`` occupying the top of the id range in reverse order (``<extra_id_0>`` = ``vocab_size-1``),
``a </s>`` / ``a </s> b </s>`` when ``add_bos_token`` (sic: the flag gates the trailing EOS).
"""
from __future__ import annotations

import logging
import os
import re
import warnings
from shutil import copyfile
from typing import List, Optional

import sentencepiece as spm

from .tokenization_base import PreTrainedTokenizer

logger = logging.getLogger(__name__)

VOCAB_FILES_NAMES = {"vocab_file": "spiece.model"}
_NAMES = ["t5-small", "t5-base", "t5-large", "t5-3b", "t5-11b"]
PRETRAINED_VOCAB_FILES_MAP = {"vocab_file": {n: f"https://huggingface.co/{n}/resolve/main/spiece.model" for n in _NAMES}}
PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES = {n: 512 for n in _NAMES}


class T5Tokenizer(PreTrainedTokenizer):
    vocab_files_names = VOCAB_FILES_NAMES
    pretrained_vocab_files_map = PRETRAINED_VOCAB_FILES_MAP
    max_model_input_sizes = PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES

    def __init__(self, vocab_file, eos_token="</s>", unk_token="<unk>", pad_token="<pad>", extra_ids=100,
                 additional_special_tokens=None, add_bos_token=False, **kwargs):
        if extra_ids > 0 and additional_special_tokens is None:
            additional_special_tokens = [f"<extra_id_{i}>" for i in range(extra_ids)]
        elif extra_ids > 0:
            n_extra = len({t for t in additional_special_tokens if "extra_id" in str(t)})
            if n_extra != extra_ids:
                raise ValueError(
                    f"Both extra_ids ({extra_ids}) and additional_special_tokens ({additional_special_tokens}) are "
                    "provided to T5Tokenizer. In this case the additional_special_tokens must include the extra_ids tokens"
                )
        super().__init__(eos_token=eos_token, unk_token=unk_token, pad_token=pad_token,
                         additional_special_tokens=additional_special_tokens, **kwargs)
        self.vocab_file = vocab_file
        self._extra_ids = extra_ids
        self.sp_model = spm.SentencePieceProcessor()
        self.sp_model.Load(vocab_file)
        self.add_bos_token = add_bos_token

    @property
    def vocab_size(self):
        return self.sp_model.get_piece_size() + self._extra_ids

    def get_vocab(self):
        vocab = {self.convert_ids_to_tokens(i): i for i in range(self.vocab_size)}
        vocab.update(self.added_tokens_encoder)
        return vocab

    def _tokenize(self, text):
        return self.sp_model.encode(text, out_type=str)

    def _convert_token_to_id(self, token):
        m = re.match(r"<extra_id_(\d+)>", token)
        if m:
            return self.vocab_size - int(m.group(1)) - 1
        return self.sp_model.piece_to_id(token)

    def _convert_id_to_token(self, index):
        if index < self.sp_model.get_piece_size():
            return self.sp_model.IdToPiece(index)
        return f"<extra_id_{self.vocab_size - 1 - index}>"

    def convert_tokens_to_string(self, tokens):
        special = set(self.all_special_tokens)
        out, current = "", []
        for token in tokens:
            if token in special:
                out += self.sp_model.decode_pieces(current) + token + " "
                current = []
            else:
                current.append(token)
        return (out + self.sp_model.decode_pieces(current)).strip()

    def _add_eos_if_not_present(self, token_ids):
        if not self.add_bos_token:
            return token_ids
        if len(token_ids) > 0 and token_ids[-1] == self.eos_token_id:
            warnings.warn(f"This sequence already has {self.eos_token}.")
            return token_ids
        return token_ids + [self.eos_token_id]

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        token_ids_0 = self._add_eos_if_not_present(token_ids_0)
        if token_ids_1 is None:
            return token_ids_0
        return token_ids_0 + self._add_eos_if_not_present(token_ids_1)

    def save_vocabulary(self, save_directory, filename_prefix=None):
        if not os.path.isdir(save_directory):
            logger.error(f"Vocabulary path ({save_directory}) should be a directory")
            return None
        out = os.path.join(save_directory, (filename_prefix + "-" if filename_prefix else "") + VOCAB_FILES_NAMES["vocab_file"])
        if os.path.abspath(self.vocab_file) != os.path.abspath(out):
            copyfile(self.vocab_file, out)
        return (out,)
