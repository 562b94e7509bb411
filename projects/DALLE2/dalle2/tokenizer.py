"""CLIP byte-pair tokenizer front end (reference projects/DALLE2/dalle2/tokenizer.py): ``tokenize(texts) →
LongTensor[B, 77]`` zero-padded, truncated keeping the EOT token."""
import torch

from projects.CLIP.clip.simple_tokenizer import SimpleTokenizer as _BPE


class SimpleTokenizer(_BPE):
    def tokenize(self, texts, context_length=77, truncate_text=True):
        if isinstance(texts, str):
            texts = [texts]
        sot, eot = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, text in enumerate(texts):
            ids = [sot] + self.encode(text) + [eot]
            if len(ids) > context_length:
                if not truncate_text:
                    raise RuntimeError(f"Input {text!r} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = eot
            out[i, : len(ids)] = torch.tensor(ids)
        return out


class YttmTokenizer:
    """YouTokenToMe BPE front-end with the same ``encode / decode / tokenize`` surface as ``SimpleTokenizer`` (reference
    projects/DALLE2/dalle2/tokenizer.py:179-219).  ``youtokentome`` is an optional dependency: the import happens in the
    constructor so the rest of the project works without it."""

    def __init__(self, bpe_path=None):
        import os

        assert bpe_path is not None and os.path.exists(str(bpe_path)), f"BPE model path {bpe_path} does not exist"
        try:
            import youtokentome as yttm
        except ImportError as e:  # pragma: no cover - optional dependency
            raise ImportError("YttmTokenizer needs `youtokentome` (pip install youtokentome)") from e
        self.yttm = yttm
        self.tokenizer = yttm.BPE(model=str(bpe_path))
        self.vocab_size = self.tokenizer.vocab_size()

    def decode(self, tokens, pad_tokens=frozenset()):
        import torch

        if torch.is_tensor(tokens):
            tokens = tokens.tolist()
        return self.tokenizer.decode(tokens, ignore_ids=set(pad_tokens) | {0})

    def encode(self, texts):
        import torch

        return [torch.tensor(ids, dtype=torch.long) for ids in self.tokenizer.encode(texts, output_type=self.yttm.OutputType.ID)]

    def tokenize(self, texts, context_length=256, truncate_text=False):
        import torch

        texts = [texts] if isinstance(texts, str) else list(texts)
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, ids in enumerate(self.encode(texts)):
            if len(ids) > context_length:
                if not truncate_text:
                    raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
                ids = ids[:context_length]
            out[i, : len(ids)] = ids
        return out
