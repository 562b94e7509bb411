"""Import-path shim: the reference keeps this class in its own file (projects/MT5/layers/mlp_layer.py); the implementation lives in projects/MT5/mt5_model.py."""
from projects.MT5.mt5_model import MT5MLP, T5MLP  # noqa: F401
