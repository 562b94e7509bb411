"""BLOOM uses the tanh GELU approximation (reference projects/BLOOM/modeling/activation.py); it runs inside the
GEMM epilogue (``Linear(..., act="gelu_tanh")``)."""
from libai_b200.ops.functional import gelu_ref  # noqa: F401


def bloom_gelu_forward(x):
    return gelu_ref(x, approximate="tanh")
