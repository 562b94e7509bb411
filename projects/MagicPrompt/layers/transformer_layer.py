"""Import-path shim: the reference keeps this class in its own file (projects/MagicPrompt/layers/transformer_layer.py); the implementation lives in libai_b200/layers/transformer_layer.py (KV cache built in)."""
from libai_b200.layers import TransformerLayer  # noqa: F401
