"""Dropout module backed by the native Philox kernel.

Reference: ``flow.nn.Dropout`` in the embeddings (libai/models/gpt_model.py:131, bert_model.py:93).  ``sharded`` says
whether the input is sharded over the tensor-parallel group (different mask per rank) or replicated (identical mask on
all ranks) — see :func:`libai_b200.ops.functional.tp_rng_salt`."""
from torch import nn

from libai_b200.ops import functional as OF


class Dropout(nn.Module):
    def __init__(self, p: float = 0.5, sharded: bool = False):
        super().__init__()
        if p < 0 or p >= 1.0 and p != 1.0:
            raise ValueError(f"dropout probability has to be in [0, 1], got {p}")
        self.p = float(p)
        self.sharded = sharded

    def forward(self, x):
        if self.p >= 1.0 and self.training:
            return x * 0
        return OF.dropout(x, self.p, self.training, sharded=self.sharded)

    def extra_repr(self) -> str:
        return f"p={self.p}, sharded={self.sharded}"
