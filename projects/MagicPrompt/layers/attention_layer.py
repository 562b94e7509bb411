"""Import-path shim: the reference keeps this class in its own file (projects/MagicPrompt/layers/attention_layer.py); the implementation lives in libai_b200/layers/attention.py (KV cache built in)."""
from libai_b200.layers import MultiheadAttention  # noqa: F401
