"""GPT-2 byte-level BPE tokenizer.

Spec: reference libai/tokenizer/tokenization_gpt2.py — ``bytes_to_unicode`` (:47-71), ``get_pairs`` (:74-85),
``GPT2Tokenizer`` (:88-272): ``vocab.json`` + ``merges.txt``, GPT-2 pre-tokenisation regex, cached greedy merges by
rank, optional leading ``<|endoftext|>`` (``add_bos_token``).  The byte-BPE core is shared with the RoBERTa
tokenizer (``ByteLevelBPE``).
"""
from __future__ import annotations

import json
import logging
import os
from functools import lru_cache
from typing import Dict, List, Optional, Tuple

import regex as re

from .tokenization_base import PreTrainedTokenizer

logger = logging.getLogger(__name__)

VOCAB_FILES_NAMES = {"vocab_file": "vocab.json", "merges_file": "merges.txt"}
PRETRAINED_VOCAB_FILES_MAP = {
    "vocab_file": {"gpt2": "https://huggingface.co/gpt2/resolve/main/vocab.json"},
    "merges_file": {"gpt2": "https://huggingface.co/gpt2/resolve/main/merges.txt"},
}
PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES = {"gpt2": 1024}

_PRETOKENIZE = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """Reversible byte → printable-unicode-char table: printable latin-1 bytes map to themselves, the other 68
    bytes are shifted to code points 256+ so that no vocabulary entry contains whitespace/control characters."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def get_pairs(word: Tuple[str, ...]):
    """Set of adjacent symbol pairs of a word given as a tuple of symbols."""
    return set(zip(word[:-1], word[1:]))


class ByteLevelBPE:
    """vocab.json / merges.txt holder with the cached merge loop."""

    def __init__(self, vocab_file, merges_file, errors="replace"):
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder: Dict[str, int] = json.load(f)
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.errors = errors
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(ln.split()) for ln in lines[1:] if ln.strip()]  # first line: "#version"
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.cache: Dict[str, str] = {}
        self.pat = re.compile(_PRETOKENIZE)

    def bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token)
        while len(word) > 1:
            pairs = get_pairs(word)
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def tokenize(self, text: str) -> List[str]:
        out = []
        for piece in re.findall(self.pat, text):
            mapped = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            out.extend(self.bpe(mapped).split(" "))
        return out

    def detokenize(self, tokens: List[str]) -> str:
        return bytearray(self.byte_decoder[c] for c in "".join(tokens)).decode("utf-8", errors=self.errors)

    def save(self, save_directory, filename_prefix=None):
        if not os.path.isdir(save_directory):
            logger.error(f"Vocabulary path ({save_directory}) should be a directory")
            return None
        prefix = filename_prefix + "-" if filename_prefix else ""
        vocab_file = os.path.join(save_directory, prefix + VOCAB_FILES_NAMES["vocab_file"])
        merge_file = os.path.join(save_directory, prefix + VOCAB_FILES_NAMES["merges_file"])
        with open(vocab_file, "w", encoding="utf-8") as f:
            f.write(json.dumps(self.encoder, ensure_ascii=False))
        with open(merge_file, "w", encoding="utf-8") as w:
            w.write("#version: 0.2\n")
            for i, (pair, rank) in enumerate(sorted(self.bpe_ranks.items(), key=lambda kv: kv[1])):
                if i != rank:
                    logger.warning(f"Saving vocabulary to {merge_file}: BPE merge indices are not consecutive.")
                w.write(" ".join(pair) + "\n")
        return vocab_file, merge_file


class _ByteBPETokenizer(PreTrainedTokenizer):
    """Shared plumbing of the GPT-2 and RoBERTa tokenizers."""

    def _init_bpe(self, vocab_file, merges_file, errors):
        self._bpe = ByteLevelBPE(vocab_file, merges_file, errors)
        self.encoder, self.decoder = self._bpe.encoder, self._bpe.decoder
        self.byte_encoder, self.byte_decoder = self._bpe.byte_encoder, self._bpe.byte_decoder
        self.bpe_ranks, self.cache, self.pat, self.errors = self._bpe.bpe_ranks, self._bpe.cache, self._bpe.pat, errors

    @property
    def vocab_size(self):
        return len(self.encoder)

    def get_vocab(self):
        return dict(self.encoder, **self.added_tokens_encoder)

    def bpe(self, token):
        return self._bpe.bpe(token)

    def _tokenize(self, text):
        return self._bpe.tokenize(text)

    def _convert_token_to_id(self, token):
        return self.encoder.get(token, self.encoder.get(self.unk_token))

    def _convert_id_to_token(self, index):
        return self.decoder.get(index)

    def convert_tokens_to_string(self, tokens):
        return self._bpe.detokenize(tokens)

    def save_vocabulary(self, save_directory, filename_prefix=None):
        return self._bpe.save(save_directory, filename_prefix)


class GPT2Tokenizer(_ByteBPETokenizer):
    vocab_files_names = VOCAB_FILES_NAMES
    pretrained_vocab_files_map = PRETRAINED_VOCAB_FILES_MAP
    max_model_input_sizes = PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES

    def __init__(self, vocab_file, merges_file, errors="replace", unk_token="<|endoftext|>",
                 bos_token="<|endoftext|>", eos_token="<|endoftext|>", add_bos_token=False, **kwargs):
        super().__init__(bos_token=bos_token, eos_token=eos_token, unk_token=unk_token, **kwargs)
        self._init_bpe(vocab_file, merges_file, errors)
        self.add_bos_token = add_bos_token

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        """``<|endoftext|> a`` / ``<|endoftext|> a <|endoftext|> b`` when ``add_bos_token``."""
        bos = [self.bos_token_id] if self.add_bos_token else []
        if token_ids_1 is None:
            return bos + token_ids_0
        return bos + token_ids_0 + bos + token_ids_1
