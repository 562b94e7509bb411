import torch
from torch import nn

from libai_b200.layers import Linear


def cosine_similarity(x, y, dim=-1):
    return torch.nn.functional.cosine_similarity(x.float(), y.float(), dim=dim)


class MLPLayer(nn.Module):
    """Projection head used only at training time (reference projects/SimCSE/modeling/model_utils.py:11-22)."""

    def __init__(self, cfg):
        super().__init__()
        self.dense = Linear(cfg.hidden_size, cfg.hidden_size, bias=True, parallel="data", layer_idx=-1)
        self.activation = nn.Tanh()

    def forward(self, features):
        return self.activation(self.dense(features))
