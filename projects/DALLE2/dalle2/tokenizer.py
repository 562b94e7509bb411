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
