"""ChatGLM3 tokenizer (reference projects/ChatGLM/tokenizer.py): sentencepiece model plus the control tokens
``[MASK] [gMASK] [sMASK] sop eop <|system|> <|user|> <|assistant|> <|observation|>`` appended after the piece ids,
``[gMASK] sop`` prefix, and the chat template builder."""
import json
import re
from enum import Enum
from typing import Dict, List, Optional

import torch


class PaddingStrategy(str, Enum):
    """Values of the ``padding_strategy`` argument of :meth:`ChatGLMTokenizer._pad`."""

    LONGEST = "longest"
    MAX_LENGTH = "max_length"


class SPTokenizer:
    def __init__(self, model_path: str):
        from sentencepiece import SentencePieceProcessor

        self.sp_model = SentencePieceProcessor(model_file=model_path)
        self.n_words = self.sp_model.vocab_size()
        self.bos_id, self.eos_id = self.sp_model.bos_id(), self.sp_model.eos_id()
        self.pad_id = self.sp_model.unk_id()
        role_tokens = ["<|system|>", "<|user|>", "<|assistant|>", "<|observation|>"]
        specials = ["[MASK]", "[gMASK]", "[sMASK]", "sop", "eop"] + role_tokens
        self.special_tokens: Dict[str, int] = {}
        self.index_special_tokens: Dict[int, str] = {}
        for tok in specials:
            self.special_tokens[tok] = self.n_words
            self.index_special_tokens[self.n_words] = tok
            self.n_words += 1
        self.role_special_token_expression = "|".join(re.escape(t) for t in role_tokens)

    def tokenize(self, s: str, encode_special_tokens=False):
        if not encode_special_tokens:
            return self.sp_model.EncodeAsPieces(s)
        last, out = 0, []
        for m in re.finditer(self.role_special_token_expression, s):
            if last < m.start():
                out.extend(self.sp_model.EncodeAsPieces(s[last : m.start()]))
            out.append(s[m.start() : m.end()])
            last = m.end()
        if last < len(s):
            out.extend(self.sp_model.EncodeAsPieces(s[last:]))
        return out

    def encode(self, s: str, bos: bool = False, eos: bool = False) -> List[int]:
        t = self.sp_model.encode(s)
        return ([self.bos_id] if bos else []) + t + ([self.eos_id] if eos else [])

    def decode(self, t: List[int]) -> str:
        text, buf = "", []
        for tok in t:
            if tok in self.index_special_tokens:
                if buf:
                    text += self.sp_model.decode(buf)
                    buf = []
                text += self.index_special_tokens[tok]
            else:
                buf.append(tok)
        return text + (self.sp_model.decode(buf) if buf else "")

    def convert_token_to_id(self, token):
        return self.special_tokens.get(token, self.sp_model.PieceToId(token))

    def convert_id_to_token(self, index):
        if index in self.index_special_tokens:
            return self.index_special_tokens[index]
        if index in (self.eos_id, self.bos_id, self.pad_id) or index < 0 or index > self.sp_model.vocab_size():
            return ""
        return self.sp_model.IdToPiece(index)


class ChatGLMTokenizer:
    def __init__(self, vocab_file, padding_side="left", encode_special_tokens=False, **kwargs):
        self.name = "GLMTokenizer"
        self.vocab_file = vocab_file
        self.tokenizer = SPTokenizer(vocab_file)
        self.special_tokens = {"<bos>": self.tokenizer.bos_id, "<eos>": self.tokenizer.eos_id,
                               "<unk>": self.tokenizer.pad_id, "<pad>": self.tokenizer.pad_id}
        self.encode_special_tokens, self.padding_side = encode_special_tokens, padding_side
        self.eod_token = None

    def get_command(self, token):
        if token in self.special_tokens:
            return self.special_tokens[token]
        assert token in self.tokenizer.special_tokens, f"{token} is not a special token for {self.name}"
        return self.tokenizer.special_tokens[token]

    @property
    def pad_token_id(self):
        return self.get_command("<pad>")

    @property
    def eos_token_id(self):
        return self.get_command("<eos>")

    @property
    def vocab_size(self):
        return self.tokenizer.n_words

    def __len__(self):
        return self.vocab_size

    def padded_vocab_size(self, multiple=1):
        return (self.vocab_size + multiple - 1) // multiple * multiple

    def get_prefix_tokens(self):
        return [self.get_command("[gMASK]"), self.get_command("sop")]

    def encode(self, text, add_special_tokens=True, return_tensors=None, **kwargs):
        ids = self.tokenizer.encode(text)
        if add_special_tokens:
            ids = self.get_prefix_tokens() + ids
        return torch.tensor([ids], dtype=torch.long) if return_tensors in ("pt", "of") else ids

    def tokenize(self, text, add_bos=True, add_eos=False, padding=False, device=None, max_length=4096, **kwargs):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [(self.get_prefix_tokens() if add_bos else []) + self.tokenizer.encode(t)[:max_length] +
                ([self.eos_token_id] if add_eos else []) for t in texts]
        out = torch.tensor(self._pad(rows), dtype=torch.long)
        return out.to(device) if device and (device != "cuda" or torch.cuda.is_available()) else out

    def _pad(self, encoded_inputs: List[List[int]], max_length: Optional[int] = None,
             padding_strategy: PaddingStrategy = PaddingStrategy.LONGEST) -> List[List[int]]:
        """Pad a batch of id lists on ``padding_side`` to the longest row (or to ``max_length``)."""
        if PaddingStrategy(padding_strategy) == PaddingStrategy.LONGEST or max_length is None:
            max_length = max(len(r) for r in encoded_inputs)
        pad = self.pad_token_id
        if self.padding_side == "left":
            return [[pad] * (max_length - len(r)) + list(r) for r in encoded_inputs]
        return [list(r) + [pad] * (max_length - len(r)) for r in encoded_inputs]

    def decode(self, ids, skip_special_tokens=True, **kwargs):
        if torch.is_tensor(ids):
            ids = ids.tolist()
        return self.tokenizer.decode(ids)

    def build_single_message(self, role, metadata, message):
        assert role in ["system", "user", "assistant", "observation"], role
        role_tokens = [self.get_command(f"<|{role}|>")] + self.tokenizer.encode(f"{metadata}\n")
        return role_tokens + self.tokenizer.encode(message)

    def build_chat_input(self, query, history: Optional[List[dict]] = None, role="user"):
        ids = []
        for item in history or []:
            content = item["content"]
            if item["role"] == "system" and "tools" in item:
                content = content + "\n" + json.dumps(item["tools"], indent=4, ensure_ascii=False)
            ids.extend(self.build_single_message(item["role"], item.get("metadata", ""), content))
        ids.extend(self.build_single_message(role, "", query))
        ids.extend([self.get_command("<|assistant|>")])
        ids = self.get_prefix_tokens() + ids
        return {"input_ids": torch.tensor([ids], dtype=torch.long)}
