"""Vocabulary-parallel cross entropy.

Spec: reference libai/layers/cross_entropy.py:21-48 — logits ``[b, s, V/t]`` split over the TP
group, integer targets ``[b, s]``; negative targets are clamped to 0 (the caller masks them);
returns the *per-token* loss ``[b, s]`` (fp32).
"""
from torch import nn

from libai_b200.ops import functional as OF
from libai_b200.utils import distributed as dutil


class ParallelCrossEntropyLoss(nn.Module):
    def forward(self, logits, target):
        assert logits.ndim == 3 and target.ndim == 2 and logits.shape[:2] == target.shape
        topo = dutil.get_dist_util()
        target = target * (target >= 0)
        start = topo.tp_rank * logits.shape[-1] if topo.tensor_parallel_size > 1 else 0
        return OF.vocab_parallel_cross_entropy(logits, target, start, topo.tp_group)
