"""Initialisers (spec: reference libai/models/utils/weight_init.py:21-37): ``N(0, σ)`` and the
residual-scaled ``N(0, σ/√(2L))`` used for output projections.  Both accept a ``generator`` so
parameter values are independent of the parallel layout (see layers/_param.py)."""
import math


def init_method_normal(sigma, mean=0.0):
    def init_(tensor, generator=None):
        return tensor.normal_(mean, sigma, generator=generator)

    return init_


def scaled_init_method_normal(sigma, num_layers, mean=0.0):
    std = sigma / math.sqrt(2.0 * num_layers)

    def init_(tensor, generator=None):
        return tensor.normal_(mean, std, generator=generator)

    return init_
