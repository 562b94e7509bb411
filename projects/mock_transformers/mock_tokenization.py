"""HF tokenizers whose batches land where the tensor-parallel model expects them.

Spec: reference projects/mock_transformers/mock_tokenization.py — there the HF tokenizers are taught a new
``return_tensors="of"`` type producing (optionally global) oneflow tensors.  Here ``"pt"`` already is the native
type; what is left is the *placement*: every tensor of the returned ``BatchEncoding`` is moved to this rank's
device (and, for ``is_global=True``, broadcast from the first rank so all tensor-parallel ranks decode the same
batch even when their tokenizer inputs differ).
"""
from __future__ import annotations

import torch

from libai_b200.utils import distributed as dist

__all__ = ["wrap_tokenizer", "BertTokenizer", "GPT2Tokenizer", "T5Tokenizer", "MT5Tokenizer", "Qwen2Tokenizer"]


def _place(batch, is_global: bool):
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    for k, v in list(batch.items()):
        if isinstance(v, torch.Tensor):
            v = v.to(device)
            if is_global and torch.distributed.is_initialized() and dist.get_world_size() > 1:
                torch.distributed.broadcast(v, src=0)
            batch[k] = v
    return batch


def wrap_tokenizer(tokenizer):
    """Return ``tokenizer`` with ``__call__``/``batch_encode_plus`` accepting ``return_tensors="of"`` (alias of
    ``"pt"`` + device placement) and ``is_global=True``."""
    cls = tokenizer.__class__

    class Placed(cls):
        def __call__(self, *args, return_tensors=None, is_global=False, **kwargs):
            placed = return_tensors == "of"
            out = super().__call__(*args, return_tensors="pt" if placed else return_tensors, **kwargs)
            return _place(out, is_global) if placed else out

    Placed.__name__ = cls.__name__
    tokenizer.__class__ = Placed
    return tokenizer


def _lazy(name):
    def from_pretrained(*args, **kwargs):
        import transformers

        return wrap_tokenizer(getattr(transformers, name).from_pretrained(*args, **kwargs))

    return type(name, (), {"from_pretrained": staticmethod(from_pretrained)})


BertTokenizer = _lazy("BertTokenizer")
GPT2Tokenizer = _lazy("GPT2Tokenizer")
T5Tokenizer = _lazy("T5Tokenizer")
MT5Tokenizer = _lazy("MT5Tokenizer")
Qwen2Tokenizer = _lazy("Qwen2Tokenizer")
