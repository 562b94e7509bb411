"""Tied-embedding LM head.

Spec: reference libai/layers/lm_logits.py:22-61 — ``logits = x · Eᵀ`` with the (vocab-split) word
embedding matrix, optional vocab-split bias; output stays split over the vocabulary
(consumed by ``ParallelCrossEntropyLoss``).
"""
from torch import nn

from libai_b200.ops import functional as OF
from libai_b200.parallel import mappings
from libai_b200.utils import distributed as dutil

from ._param import create_parameter, zeros_


class LMLogits(nn.Module):
    def __init__(self, vocab_size, bias=False):
        super().__init__()
        self.bias = create_parameter((vocab_size,), zeros_, tp_dim=0, layer_idx=-1) if bias else None

    def forward(self, input, word_embeddings):
        """``word_embeddings`` is the local ``[V/t, h]`` shard (column-parallel semantics)."""
        topo = dutil.get_dist_util()
        x = mappings.gather_from_sp(input) if topo.sequence_parallel else mappings.copy_to_tp(input)
        return OF.linear(x, word_embeddings, self.bias)
