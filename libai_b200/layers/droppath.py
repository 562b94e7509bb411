"""Stochastic depth (spec: reference libai/layers/droppath.py:20-46)."""
import torch
from torch import nn


def drop_path(x, drop_prob: float = 0.5, training: bool = False, scale_by_keep: bool = True):
    """Zero whole samples (dim 0) of the residual branch with probability ``drop_prob``."""
    if drop_prob == 0.0 or not training:
        return x
    keep = 1.0 - drop_prob
    mask_shape = (x.shape[0],) + (1,) * (x.dim() - 1)
    mask = torch.empty(mask_shape, dtype=x.dtype, device=x.device).bernoulli_(keep)
    if keep > 0.0 and scale_by_keep:
        mask.div_(keep)
    return x * mask


class DropPath(nn.Module):
    def __init__(self, drop_prob=None, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob or 0.0
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training, self.scale_by_keep)

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}"
