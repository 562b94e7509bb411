from .build import _expand_param_groups, build_optimizer, get_default_optimizer_params, reduce_param_groups
from .grad_scaler import DynamicLossScaler
from .build import set_weight_decay
from .optimizers import LAMB, SGD, Adam, AdamW, FlatOptimizer

__all__ = [
    "build_optimizer", "get_default_optimizer_params", "reduce_param_groups", "AdamW", "Adam", "SGD", "LAMB", "set_weight_decay",
    "FlatOptimizer", "DynamicLossScaler",
]
