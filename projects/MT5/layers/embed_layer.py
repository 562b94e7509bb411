"""Import-path shim: the reference keeps this class in its own file (projects/MT5/layers/embed_layer.py); the implementation lives in projects/MT5/mt5_model.py."""
from projects.MT5.mt5_model import MT5Embedding  # noqa: F401

Embedding = MT5Embedding
