"""BLOOM uses the tanh GELU approximation (reference projects/BLOOM/modeling/activation.py); it runs inside the
GEMM epilogue (``Linear(..., act="gelu_tanh")``)."""
from libai_b200.ops.functional import gelu_ref  # noqa: F401


def bloom_gelu_forward(x):
    return gelu_ref(x, approximate="tanh")


def bloom_gelu_back(g, x):
    """Gradient of the tanh GELU: ``g · d/dx [0.5 x (1 + tanh(√(2/π)(x + 0.044715 x³)))]``."""
    import torch

    t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
    ff = 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)
    return ff * g


class GeLUFunction:
    """Autograd function of the tanh GELU with the closed-form backward (kept for API parity; ``apply`` routes through
    ``torch.autograd.Function``)."""

    @staticmethod
    def apply(x):
        return _GeLUFn.apply(x)


import torch as _torch  # noqa: E402


class _GeLUFn(_torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return bloom_gelu_forward(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return bloom_gelu_back(g, x)


class BloomGelu(_torch.nn.Module):
    """Module form: closed-form-gradient autograd function in training, plain expression in eval."""

    def forward(self, x):
        return GeLUFunction.apply(x) if self.training else bloom_gelu_forward(x)
