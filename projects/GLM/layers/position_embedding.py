"""Import-path shim: the reference keeps this class in its own file (projects/GLM/layers/position_embedding.py); the implementation lives in libai_b200/layers/embedding.py."""
from libai_b200.layers import SinePositionalEmbedding  # noqa: F401
