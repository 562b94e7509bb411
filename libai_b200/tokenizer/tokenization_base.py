"""Tokenizer base class.

Spec: reference libai/tokenizer/tokenization_base.py — ``PreTrainedTokenizer`` (:88-1289): special-token
attributes incl. LiBai's ``eod_token`` and the ``start_token``/``end_token`` aliases (:1032-1069), added-token
tables on top of any base vocabulary (:521-676), no-split tokenisation (:678-751), ``encode`` with
``return_tensors`` (:777-843; here ``"pt"``/``"of"`` → ``torch.long`` tensor, ``"np"`` → int64 array),
``decode`` (:892-946), ``from_pretrained``/``save_pretrained`` file layout (``added_tokens.json``,
``special_tokens_map.json``, ``tokenizer_config.json``; :194-483), ``padded_vocab_size(multiple)`` (:489-494).
Written from the behaviour, not the text, of the reference.
"""
from __future__ import annotations

import copy
import itertools
import json
import logging
import os
import unicodedata
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

logger = logging.getLogger(__name__)

SPECIAL_TOKENS_MAP_FILE = "special_tokens_map.json"
ADDED_TOKENS_FILE = "added_tokens.json"
TOKENIZER_CONFIG_FILE = "tokenizer_config.json"


def _is_whitespace(char: str) -> bool:
    if char in (" ", "\t", "\n", "\r"):
        return True
    return unicodedata.category(char) == "Zs"


def _is_control(char: str) -> bool:
    if char in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(char).startswith("C")


def _is_punctuation(char: str) -> bool:
    cp = ord(char)
    # all non-alphanumeric ASCII counts as punctuation (so "$", "^", "`" split like in BERT)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(char).startswith("P")


class PreTrainedTokenizer:
    vocab_files_names: Dict[str, str] = {}
    pretrained_vocab_files_map: Dict[str, Dict[str, str]] = {}
    pretrained_init_configuration: Dict[str, dict] = {}
    max_model_input_sizes: Dict[str, Optional[int]] = {}

    SPECIAL_TOKENS_ATTRIBUTES = [
        "bos_token", "eos_token", "unk_token", "sep_token", "pad_token", "cls_token", "mask_token", "eod_token",
        "additional_special_tokens",
    ]

    def __init__(self, verbose=True, **kwargs):
        for attr in self.SPECIAL_TOKENS_ATTRIBUTES[:-1]:
            setattr(self, "_" + attr, None)
        self._additional_special_tokens: List[str] = []
        self.verbose = verbose
        self.added_tokens_encoder: Dict[str, int] = {}
        self.added_tokens_decoder: Dict[int, str] = {}
        self.unique_no_split_tokens: List[str] = []
        self.init_inputs = ()
        self.init_kwargs: dict = {}
        for key, value in kwargs.items():
            if value is None or key not in self.SPECIAL_TOKENS_ATTRIBUTES:
                continue
            if key == "additional_special_tokens":
                assert all(isinstance(t, str) for t in value), "One of the tokens is not a string"
                setattr(self, key, list(value))
            elif isinstance(value, str):
                setattr(self, key, value)
            else:
                raise TypeError(f"special token {key} has to be str but got: {type(value)}")

    # ------------------------------------------------------------------ (de)serialisation
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *init_inputs, **kwargs):
        """Load from a directory written by :meth:`save_pretrained` (or holding the files named in
        ``vocab_files_names``), from a single vocabulary file, or from a known short-cut name whose files are
        fetched through ``libai_b200.utils.file_utils.cached_path`` (needs network / a warm cache)."""
        name = str(pretrained_model_name_or_path)
        files: Dict[str, Optional[str]] = {}
        init_configuration: dict = {}
        known = set(itertools.chain.from_iterable(m.keys() for m in cls.pretrained_vocab_files_map.values()))
        if name in known and not os.path.exists(name):
            from libai_b200.utils.file_utils import cached_path

            for file_id, table in cls.pretrained_vocab_files_map.items():
                files[file_id] = cached_path(table[name], cache_dir=kwargs.get("cache_dir"))
            init_configuration = dict(cls.pretrained_init_configuration.get(name, {}))
        elif os.path.isdir(name):
            extra = {
                "added_tokens_file": ADDED_TOKENS_FILE,
                "special_tokens_map_file": SPECIAL_TOKENS_MAP_FILE,
                "tokenizer_config_file": TOKENIZER_CONFIG_FILE,
            }
            for file_id, file_name in {**cls.vocab_files_names, **extra}.items():
                full = os.path.join(name, file_name)
                files[file_id] = full if os.path.exists(full) else None
        elif os.path.isfile(name):
            if len(cls.vocab_files_names) != 1:
                raise ValueError(
                    f"{cls.__name__} needs {sorted(cls.vocab_files_names)}; pass a directory instead of a single file"
                )
            files[next(iter(cls.vocab_files_names))] = name
        else:
            raise EnvironmentError(
                f"Can't load tokenizer '{name}': not a local path and not one of {sorted(known)}"
            )
        missing = [k for k in cls.vocab_files_names if files.get(k) is None]
        if missing:
            raise EnvironmentError(f"Can't load tokenizer '{name}': missing vocabulary files {missing}")

        cfg_file = files.pop("tokenizer_config_file", None)
        init_kwargs = dict(init_configuration)
        if cfg_file is not None:
            with open(cfg_file, encoding="utf-8") as f:
                init_kwargs.update(json.load(f))
            saved_inputs = init_kwargs.pop("init_inputs", ())
            if not init_inputs:
                init_inputs = tuple(saved_inputs)
        init_kwargs.update({k: v for k, v in kwargs.items() if k != "cache_dir"})
        if name in cls.max_model_input_sizes and cls.max_model_input_sizes[name] is not None:
            init_kwargs.setdefault("max_len", cls.max_model_input_sizes[name])
        added_tokens_file = files.pop("added_tokens_file", None)
        special_tokens_map_file = files.pop("special_tokens_map_file", None)
        for file_id, path in files.items():
            init_kwargs.setdefault(file_id, path)
        if special_tokens_map_file is not None:
            with open(special_tokens_map_file, encoding="utf-8") as f:
                for key, value in json.load(f).items():
                    init_kwargs.setdefault(key, value)
        max_len = init_kwargs.pop("max_len", None)
        tokenizer = cls(*init_inputs, **init_kwargs)
        tokenizer.max_len = max_len if max_len is not None else int(1e12)
        tokenizer.init_inputs = init_inputs
        tokenizer.init_kwargs = init_kwargs
        if added_tokens_file is not None:
            with open(added_tokens_file, encoding="utf-8") as f:
                added = json.load(f)
            tokenizer.added_tokens_encoder.update(added)
            tokenizer.added_tokens_decoder.update({v: k for k, v in added.items()})
            tokenizer.unique_no_split_tokens = sorted(set(tokenizer.unique_no_split_tokens) | set(added))
        tokenizer.sanitize_special_tokens()
        return tokenizer

    def save_pretrained(self, save_directory):
        if not os.path.isdir(save_directory):
            os.makedirs(save_directory, exist_ok=True)
        cfg = copy.deepcopy(self.init_kwargs)
        if self.init_inputs:
            cfg["init_inputs"] = list(self.init_inputs)
        for file_id in self.vocab_files_names:
            cfg.pop(file_id, None)
        with open(os.path.join(save_directory, TOKENIZER_CONFIG_FILE), "w", encoding="utf-8") as f:
            f.write(json.dumps(cfg, ensure_ascii=False))
        with open(os.path.join(save_directory, SPECIAL_TOKENS_MAP_FILE), "w", encoding="utf-8") as f:
            f.write(json.dumps(self.special_tokens_map, ensure_ascii=False))
        if self.added_tokens_encoder:
            with open(os.path.join(save_directory, ADDED_TOKENS_FILE), "w", encoding="utf-8") as f:
                f.write(json.dumps(self.added_tokens_encoder, ensure_ascii=False))
        vocab_files = self.save_vocabulary(save_directory)
        return tuple(vocab_files) + (
            os.path.join(save_directory, SPECIAL_TOKENS_MAP_FILE), os.path.join(save_directory, ADDED_TOKENS_FILE),
        )

    def save_vocabulary(self, save_directory):
        raise NotImplementedError

    # ------------------------------------------------------------------ vocabulary
    @property
    def vocab_size(self) -> int:
        raise NotImplementedError

    def padded_vocab_size(self, multiple=1) -> int:
        """Vocabulary size rounded up so embedding tables split evenly (e.g. 128 × tensor-parallel size)."""
        n = len(self)
        return ((n + multiple - 1) // multiple) * multiple

    def __len__(self):
        return self.vocab_size + len(self.added_tokens_encoder)

    def get_vocab(self) -> Dict[str, int]:
        raise NotImplementedError

    def get_added_vocab(self) -> Dict[str, int]:
        return self.added_tokens_encoder

    def add_tokens(self, new_tokens: Union[str, Sequence[str]], special_tokens: bool = False) -> int:
        if not new_tokens:
            return 0
        if isinstance(new_tokens, str):
            new_tokens = [new_tokens]
        to_add = []
        for token in new_tokens:
            assert isinstance(token, str), f"Token {token} has to be of type string, but got {type(token)}."
            if not special_tokens and getattr(self, "do_lower_case", False):
                token = token.lower()
            unk = self.unk_token
            if (
                token != unk
                and token not in to_add
                and (unk is None or self.convert_tokens_to_ids(token) == self.convert_tokens_to_ids(unk))
                and token not in self.added_tokens_encoder
            ):
                to_add.append(token)
        start = len(self)
        for i, token in enumerate(to_add):
            self.added_tokens_encoder[token] = start + i
            self.added_tokens_decoder[start + i] = token
        if special_tokens:
            self.unique_no_split_tokens = sorted(set(self.unique_no_split_tokens) | set(new_tokens))
        else:
            self.unique_no_split_tokens = sorted(set(self.unique_no_split_tokens) | set(to_add))
        return len(to_add)

    def sanitize_special_tokens(self) -> int:
        """Make sure every special token is in the vocabulary (adds the missing ones)."""
        return self.add_tokens(self.all_special_tokens, special_tokens=True)

    def add_special_tokens(self, special_tokens_dict: Dict[str, Union[str, List[str]]]) -> int:
        if not special_tokens_dict:
            return 0
        added = 0
        for key, value in special_tokens_dict.items():
            assert key in self.SPECIAL_TOKENS_ATTRIBUTES, f"Key {key} is not a special token"
            setattr(self, key, value)
            if key == "additional_special_tokens":
                assert isinstance(value, (list, tuple)) and all(isinstance(t, str) for t in value)
                added += self.add_tokens(value, special_tokens=True)
            else:
                assert isinstance(value, str)
                added += self.add_tokens([value], special_tokens=True)
        return added

    # ------------------------------------------------------------------ text → tokens → ids
    def tokenize(self, text: str, **kwargs) -> List[str]:
        """Split on the no-split (added / special) tokens first, run ``_tokenize`` on the rest."""
        if getattr(self, "do_lower_case", False):
            # lower-case everything except the special tokens
            import re

            escaped = [re.escape(t) for t in self.all_special_tokens]
            if escaped:
                pattern = r"(" + r"|".join(escaped) + r")|(.+?)"
                text = re.sub(pattern, lambda m: m.groups()[0] or m.groups()[1].lower(), text, flags=re.S)
            else:
                text = text.lower()
        if not text.strip():
            return []
        # special tokens are never split even when the tokenizer was constructed directly (no sanitize step)
        no_split = sorted(set(self.unique_no_split_tokens) | set(self.all_special_tokens))
        pieces = [text]
        for tok in no_split:
            nxt = []
            for piece in pieces:
                if piece in no_split:
                    nxt.append(piece)
                    continue
                parts = piece.split(tok)
                if len(parts) == 1:  # token absent: keep the text untouched (leading spaces matter for byte-BPE)
                    nxt.append(piece)
                    continue
                for i, part in enumerate(parts):
                    part = part.strip()  # the white space around a no-split token belongs to it
                    if part:
                        nxt.append(part)
                    if i < len(parts) - 1:
                        nxt.append(tok)
            pieces = nxt
        out: List[str] = []
        for piece in pieces:
            if piece in no_split:
                out.append(piece)
            else:
                out.extend(self._tokenize(piece, **kwargs))
        return out

    def _tokenize(self, text, **kwargs):
        raise NotImplementedError

    def convert_tokens_to_ids(self, tokens):
        if tokens is None:
            return None
        if isinstance(tokens, str):
            return self._convert_token_to_id_with_added_voc(tokens)
        if len(tokens) > 0 and isinstance(tokens[0], (list, tuple)):
            return [[self._convert_token_to_id_with_added_voc(t) for t in seq] for seq in tokens]
        return [self._convert_token_to_id_with_added_voc(t) for t in tokens]

    def _convert_token_to_id_with_added_voc(self, token):
        if token is None:
            return None
        if token in self.added_tokens_encoder:
            return self.added_tokens_encoder[token]
        return self._convert_token_to_id(token)

    def _convert_token_to_id(self, token):
        raise NotImplementedError

    def convert_to_tensors(self, token_ids, return_tensors=None, is_global=False, device="cuda", **kwargs):
        """``"pt"`` (alias ``"of"`` for reference configs) → ``torch.long`` tensor; ``is_global`` moves it to
        ``device`` (every rank holds the full tensor – the reference's broadcast placement); ``"np"`` → int64."""
        if return_tensors is None:
            return token_ids
        if return_tensors in ("pt", "of"):
            t = torch.tensor(token_ids, dtype=torch.long)
            if is_global and (device != "cuda" or torch.cuda.is_available()):
                t = t.to(device)
            return t
        if return_tensors == "np":
            return np.array(token_ids, dtype=np.int64)
        raise ValueError(f"unknown return_tensors={return_tensors!r}")

    def encode(self, text, return_tensors=None, is_global=False, device="cuda", **kwargs):
        build = getattr(self, "build_inputs_with_special_tokens", None)
        if isinstance(text, str):
            ids = self.convert_tokens_to_ids(self.tokenize(text))
            if build is not None:
                ids = build(ids)
            return self.convert_to_tensors(ids, return_tensors, is_global, device, **kwargs)
        if isinstance(text, (list, tuple)) and len(text) > 0 and isinstance(text[0], str):
            ids_list = [self.convert_tokens_to_ids(self.tokenize(t)) for t in text]
            if build is not None:
                ids_list = [build(ids) for ids in ids_list]
            return self.convert_to_tensors(ids_list, return_tensors, is_global, device, **kwargs)
        if isinstance(text, (list, tuple)) and len(text) > 0 and isinstance(text[0], int):
            return text
        raise ValueError(
            "Input is not valid. Should be a string, a list/tuple of strings or a list/tuple of integers."
        )

    # ------------------------------------------------------------------ ids → tokens → text
    def convert_ids_to_tokens(self, ids, skip_special_tokens: bool = False):
        if isinstance(ids, int):
            return self.added_tokens_decoder.get(ids) or self._convert_id_to_token(ids)
        special = set(self.all_special_ids) if skip_special_tokens else ()
        tokens = []
        for index in ids:
            index = int(index)
            if index in special:
                continue
            tokens.append(self.added_tokens_decoder.get(index) or self._convert_id_to_token(index))
        return tokens

    def _convert_id_to_token(self, index: int) -> str:
        raise NotImplementedError

    def convert_tokens_to_string(self, tokens: List[str]) -> str:
        return " ".join(tokens)

    def decode(self, token_ids, skip_special_tokens=False, clean_up_tokenization_spaces=True,
               spaces_between_special_tokens: bool = True):
        if isinstance(token_ids, (torch.Tensor, np.ndarray)):
            token_ids = token_ids.tolist()
        tokens = self.convert_ids_to_tokens(token_ids, skip_special_tokens=skip_special_tokens)
        # byte-level vocabularies must not be mixed with plain-text added tokens: build the string piecewise
        sub_texts, current = [], []
        for token in tokens:
            if token in self.added_tokens_encoder:
                if current:
                    sub_texts.append(self.convert_tokens_to_string(current))
                    current = []
                sub_texts.append(token)
            else:
                current.append(token)
        if current:
            sub_texts.append(self.convert_tokens_to_string(current))
        text = (" " if spaces_between_special_tokens else "").join(sub_texts)
        return self.clean_up_tokenization(text) if clean_up_tokenization_spaces else text

    @staticmethod
    def clean_up_tokenization(out_string):
        for a, b in ((" .", "."), (" ?", "?"), (" !", "!"), (" ,", ","), (" ' ", "'"), (" n't", "n't"), (" 'm", "'m"),
                     (" do not", " don't"), (" 's", "'s"), (" 've", "'ve"), (" 're", "'re")):
            out_string = out_string.replace(a, b)
        return out_string

    # ------------------------------------------------------------------ special tokens
    def _get_special(self, name):
        value = getattr(self, "_" + name)
        if value is None:
            if self.verbose:
                logger.error(f"Using {name}, but it is not set yet.")
            return None
        return str(value)

    @property
    def start_token(self) -> Optional[str]:
        """Common name for ``bos_token`` / ``cls_token``."""
        return self._alias(self._bos_token, self._cls_token, "bos_token", "cls_token", "start_token")

    @property
    def end_token(self) -> Optional[str]:
        """Common name for ``eos_token`` / ``sep_token`` (``eod_token`` is not considered)."""
        return self._alias(self._eos_token, self._sep_token, "eos_token", "sep_token", "end_token")

    def _alias(self, a, b, name_a, name_b, alias):
        if a is not None and b is not None:
            if a == b:
                return str(a)
            logger.error(f"Conflict between {name_a} and {name_b}.")
            return None
        if a is not None or b is not None:
            return str(a if a is not None else b)
        logger.error(f"Using {alias}, but it is not set yet.")
        return None

    @property
    def start_token_id(self):
        tok = self.start_token
        return None if tok is None else self.convert_tokens_to_ids(tok)

    @property
    def end_token_id(self):
        tok = self.end_token
        return None if tok is None else self.convert_tokens_to_ids(tok)

    @property
    def additional_special_tokens(self) -> List[str]:
        return [str(t) for t in self._additional_special_tokens]

    @additional_special_tokens.setter
    def additional_special_tokens(self, value):
        self._additional_special_tokens = list(value)

    @property
    def additional_special_tokens_ids(self) -> List[int]:
        return self.convert_tokens_to_ids(self.additional_special_tokens)

    @property
    def special_tokens_map(self) -> Dict[str, Union[str, List[str]]]:
        out = {}
        for attr in self.SPECIAL_TOKENS_ATTRIBUTES:
            value = getattr(self, "_" + attr)
            if value:
                out[attr] = value
        return out

    @property
    def all_special_tokens(self) -> List[str]:
        seen, out = set(), []
        for value in self.special_tokens_map.values():
            for tok in (value if isinstance(value, (list, tuple)) else [value]):
                if tok not in seen:
                    seen.add(tok)
                    out.append(str(tok))
        return out

    @property
    def all_special_ids(self) -> List[int]:
        return self.convert_tokens_to_ids(self.all_special_tokens)


def _special_property(name):
    def getter(self):
        return self._get_special(name)

    def setter(self, value):
        setattr(self, "_" + name, value)

    def id_getter(self):
        value = getattr(self, "_" + name)
        return None if value is None else self.convert_tokens_to_ids(str(value))

    return property(getter, setter), property(id_getter)


for _name in PreTrainedTokenizer.SPECIAL_TOKENS_ATTRIBUTES[:-1]:
    _prop, _id_prop = _special_property(_name)
    setattr(PreTrainedTokenizer, _name, _prop)
    setattr(PreTrainedTokenizer, _name + "_id", _id_prop)
