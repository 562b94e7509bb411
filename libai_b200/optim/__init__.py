from .build import _expand_param_groups, build_optimizer, get_default_optimizer_params, reduce_param_groups
from .grad_scaler import DynamicLossScaler
from .optimizers import SGD, Adam, AdamW, FlatOptimizer

__all__ = [
    "build_optimizer", "get_default_optimizer_params", "reduce_param_groups", "AdamW", "Adam", "SGD",
    "FlatOptimizer", "DynamicLossScaler",
]
