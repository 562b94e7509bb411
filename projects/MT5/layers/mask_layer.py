"""Import-path shim: the reference keeps this class in its own file (projects/MT5/layers/mask_layer.py); the implementation lives in projects/MT5/mt5_model.py (_extend)."""
import torch
from torch import nn

from projects.MT5.mt5_model import _extend


class ExtendedMask(nn.Module):
    """``[b, s]`` padding mask (or ``[b, q, k]``) → broadcastable attention mask; ``is_decoder`` adds the causal part."""

    def forward(self, x, is_decoder: bool = False):
        m = _extend(x)                                            # [b, 1, 1|q, k] boolean
        if is_decoder:
            k = m.shape[-1]
            m = m & torch.ones(k, k, dtype=torch.bool, device=m.device).tril()
        return m
