"""Activation registry (spec: reference libai/layers/activation.py:23-87)."""
from enum import Enum
from typing import Optional

import torch
from torch import nn


class Activation(str, Enum):
    SquaredReLU = "squared_relu"
    GeLU = "gelu"
    GeLUTanh = "gelu_tanh"
    LeakyReLU = "leaky_relu"
    ReLU = "relu"
    Tanh = "tanh"
    QuickGELU = "quick_gelu"
    SiLU = "silu"


class SquaredReLU(nn.Module):
    def forward(self, x):
        r = torch.relu(x)
        return r * r


class Passthrough(nn.Module):
    def forward(self, x):
        return x


class GeLUTanh(nn.Module):
    """tanh approximation of GELU: 0.5x(1 + tanh(√(2/π)(x + 0.044715x³)))."""

    def forward(self, x):
        return torch.nn.functional.gelu(x, approximate="tanh")


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


def build_activation(activation: Optional[Activation]):
    """Return the module for ``activation`` (``None`` → identity)."""
    if not activation:
        return Passthrough()
    table = {
        Activation.ReLU: nn.ReLU,
        Activation.GeLU: nn.GELU,
        Activation.GeLUTanh: GeLUTanh,
        Activation.LeakyReLU: nn.LeakyReLU,
        Activation.SquaredReLU: SquaredReLU,
        Activation.Tanh: nn.Tanh,
        Activation.QuickGELU: QuickGELU,
        Activation.SiLU: nn.SiLU,
    }
    return table[Activation(activation)]()
