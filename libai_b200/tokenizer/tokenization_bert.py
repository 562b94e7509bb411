"""BERT WordPiece tokenizer.

Spec: reference libai/tokenizer/tokenization_bert.py — ``BertTokenizer`` (:83-252; ``[CLS] a [SEP] b [SEP]`` only
when ``add_bos_token``), ``BasicTokenizer`` (:255-397: clean → CJK spacing → whitespace split → lower/strip accents
→ punctuation split, ``never_split``), ``BasicTokenizerWithChineseWWM`` (:400-440, jieba pre-segmentation so
non-initial characters of a Chinese word become ``##`` pieces), ``WordpieceTokenizer`` (:443-514, greedy
longest-match-first, Chinese ``##X`` pieces are looked up as ``X``).
"""
from __future__ import annotations

import collections
import logging
import os
import re
import unicodedata
from typing import List, Optional

from .tokenization_base import PreTrainedTokenizer, _is_control, _is_punctuation, _is_whitespace

logger = logging.getLogger(__name__)

VOCAB_FILES_NAMES = {"vocab_file": "vocab.txt"}
_HF = "https://huggingface.co/{}/resolve/main/vocab.txt"
_NAMES = ["bert-base-uncased", "bert-large-uncased", "bert-base-cased", "bert-large-cased", "bert-base-chinese"]
PRETRAINED_VOCAB_FILES_MAP = {"vocab_file": {n: _HF.format(n) for n in _NAMES}}
PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES = {n: 512 for n in _NAMES}
PRETRAINED_INIT_CONFIGURATION = {n: {"do_lower_case": n.endswith("uncased")} for n in _NAMES}


def load_vocab(vocab_file):
    """One token per line → ordered ``token → index``."""
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as reader:
        for index, line in enumerate(reader):
            vocab[line.rstrip("\n")] = index
    return vocab


def whitespace_tokenize(text):
    text = text.strip()
    return text.split() if text else []


def _is_chinese_substr(token):
    return re.findall("##[一-龥]", token)


def _is_cjk(cp: int) -> bool:
    return (
        0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
        or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F
    )


class BasicTokenizer:
    """Punctuation splitting, lower casing, accent stripping, CJK character isolation."""

    def __init__(self, do_lower_case=True, never_split=None, tokenize_chinese_chars=True):
        self.do_lower_case = do_lower_case
        self.never_split = list(never_split or [])
        self.tokenize_chinese_chars = tokenize_chinese_chars

    def tokenize(self, text, never_split=None):
        never_split = set(self.never_split) | set(never_split or [])
        text = self._clean_text(text)
        if self.tokenize_chinese_chars:
            text = self._tokenize_chinese_chars(text)
        out = []
        for token in whitespace_tokenize(text):
            if token in never_split:
                out.append(token)
                continue
            if self.do_lower_case:
                token = self._run_strip_accents(token.lower())
            out.extend(self._run_split_on_punc(token))
        return whitespace_tokenize(" ".join(out))

    @staticmethod
    def _run_strip_accents(text):
        return "".join(c for c in unicodedata.normalize("NFD", text) if unicodedata.category(c) != "Mn")

    @staticmethod
    def _run_split_on_punc(text):
        pieces, word = [], []
        for ch in text:
            if _is_punctuation(ch):
                if word:
                    pieces.append("".join(word))
                    word = []
                pieces.append(ch)
            else:
                word.append(ch)
        if word:
            pieces.append("".join(word))
        return pieces

    def _tokenize_chinese_chars(self, text):
        out = []
        for ch in text:
            out.append(f" {ch} " if _is_cjk(ord(ch)) else ch)
        return "".join(out)

    _is_chinese_char = staticmethod(_is_cjk)

    @staticmethod
    def _clean_text(text):
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            out.append(" " if _is_whitespace(ch) else ch)
        return "".join(out)


class BasicTokenizerWithChineseWWM(BasicTokenizer):
    """Chinese runs are segmented into words first (jieba, or a user ``pre_tokenizer``); WordPiece then marks the
    non-initial characters of a word with ``##`` so whole-word masking can find word boundaries."""

    def __init__(self, do_lower_case=True, never_split=None, tokenize_chinese_chars=True, pre_tokenizer=None):
        super().__init__(do_lower_case, never_split, tokenize_chinese_chars)
        if pre_tokenizer is None:
            try:
                import jieba
            except ImportError as e:  # pragma: no cover - depends on the environment
                raise ImportError("Chinese whole word mask need jieba (or pass pre_tokenizer=callable)") from e
            pre_tokenizer = lambda x: jieba.lcut(x, HMM=False)  # noqa: E731
        self.pre_tokenizer = pre_tokenizer

    def _tokenize_chinese_chars(self, text):
        out, run = [], []

        def flush():
            if run:
                for word in self.pre_tokenizer("".join(run)):
                    out.append(f" {word} ")
                run.clear()

        for ch in text:
            if _is_cjk(ord(ch)):
                run.append(ch)
            else:
                flush()
                out.append(ch)
        flush()
        return "".join(out)


class WordpieceTokenizer:
    def __init__(self, vocab, unk_token, max_input_chars_per_word=100):
        self.vocab, self.unk_token, self.max_input_chars_per_word = vocab, unk_token, max_input_chars_per_word

    def _known(self, piece: str) -> bool:
        if piece.startswith("##") and _is_chinese_substr(piece):
            return piece[2:] in self.vocab  # Chinese continuation pieces share the entry of the bare character(s)
        return piece in self.vocab

    def tokenize(self, text):
        out = []
        for token in whitespace_tokenize(text):
            if len(token) > self.max_input_chars_per_word:
                out.append(self.unk_token)
                continue
            start, pieces = 0, []
            while start < len(token):
                end = len(token)
                found = None
                while start < end:
                    piece = ("##" if start > 0 else "") + token[start:end]
                    if self._known(piece):
                        found = piece
                        break
                    end -= 1
                if found is None:
                    pieces = None
                    break
                pieces.append(found)
                start = end
            out.extend(pieces if pieces is not None else [self.unk_token])
        return out


class BertTokenizer(PreTrainedTokenizer):
    vocab_files_names = VOCAB_FILES_NAMES
    pretrained_vocab_files_map = PRETRAINED_VOCAB_FILES_MAP
    pretrained_init_configuration = PRETRAINED_INIT_CONFIGURATION
    max_model_input_sizes = PRETRAINED_POSITIONAL_EMBEDDINGS_SIZES

    def __init__(self, vocab_file, do_lower_case=True, do_basic_tokenize=True, never_split=None, unk_token="[UNK]",
                 sep_token="[SEP]", pad_token="[PAD]", cls_token="[CLS]", mask_token="[MASK]",
                 tokenize_chinese_chars=True, do_chinese_wwm=False, add_bos_token=False, **kwargs):
        super().__init__(unk_token=unk_token, sep_token=sep_token, pad_token=pad_token, cls_token=cls_token,
                         mask_token=mask_token, **kwargs)
        if not os.path.isfile(vocab_file):
            raise ValueError(
                f"Can't find a vocabulary file at path '{vocab_file}'. To load the vocabulary of a published model "
                "use `BertTokenizer.from_pretrained(PRETRAINED_MODEL_NAME)`"
            )
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.do_basic_tokenize = do_basic_tokenize
        self.do_chinese_wwm = do_chinese_wwm
        if do_basic_tokenize:
            cls_ = BasicTokenizerWithChineseWWM if do_chinese_wwm else BasicTokenizer
            extra = {"pre_tokenizer": kwargs["pre_tokenizer"]} if do_chinese_wwm and "pre_tokenizer" in kwargs else {}
            self.basic_tokenizer = cls_(do_lower_case=do_lower_case, never_split=never_split,
                                        tokenize_chinese_chars=tokenize_chinese_chars, **extra)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab, unk_token=self.unk_token)
        self.add_bos_token = add_bos_token

    @property
    def vocab_size(self):
        return len(self.vocab)

    def get_vocab(self):
        return dict(self.vocab, **self.added_tokens_encoder)

    def _tokenize(self, text):
        if not self.do_basic_tokenize:
            return self.wordpiece_tokenizer.tokenize(text)
        out = []
        for token in self.basic_tokenizer.tokenize(text, never_split=self.all_special_tokens):
            out.extend(self.wordpiece_tokenizer.tokenize(token))
        return out

    def _convert_token_to_id(self, token):
        """Chinese continuation pieces ``##X`` (not in the file) map to ``vocab_size + id(X)``."""
        if token in self.vocab:
            return self.vocab[token]
        if self.do_chinese_wwm and token.startswith("##") and _is_chinese_substr(token) and token[2:] in self.vocab:
            return len(self.vocab) + self.vocab[token[2:]]
        return self.vocab.get(self.unk_token)

    def _convert_id_to_token(self, index):
        if index in self.ids_to_tokens:
            return self.ids_to_tokens[index]
        if self.do_chinese_wwm and len(self.vocab) <= index < 2 * len(self.vocab):
            return "##" + self.ids_to_tokens[index - len(self.vocab)]
        return self.unk_token

    def convert_tokens_to_string(self, tokens):
        return " ".join(tokens).replace(" ##", "").strip()

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None):
        """``[CLS] a [SEP]`` / ``[CLS] a [SEP] b [SEP]`` when ``add_bos_token``; bare ids otherwise."""
        cls, sep = ([self.cls_token_id], [self.sep_token_id]) if self.add_bos_token else ([], [])
        if token_ids_1 is None:
            return cls + token_ids_0 + sep
        return cls + token_ids_0 + sep + token_ids_1 + sep

    def save_vocabulary(self, save_directory, filename_prefix=None):
        prefix = filename_prefix + "-" if filename_prefix else ""
        if os.path.isdir(save_directory):
            vocab_file = os.path.join(save_directory, prefix + VOCAB_FILES_NAMES["vocab_file"])
        else:
            vocab_file = prefix + save_directory
        expected = 0
        with open(vocab_file, "w", encoding="utf-8") as writer:
            for token, index in sorted(self.vocab.items(), key=lambda kv: kv[1]):
                if index != expected:
                    logger.warning(f"Saving vocabulary to {vocab_file}: vocabulary indices are not consecutive.")
                    expected = index
                writer.write(token + "\n")
                expected += 1
        return (vocab_file,)
