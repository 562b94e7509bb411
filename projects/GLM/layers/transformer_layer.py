"""Import-path shim: the reference keeps this class in its own file (projects/GLM/layers/transformer_layer.py); the implementation lives in projects/GLM/modeling_glm.py."""
from projects.GLM.modeling_glm import GLMLayer as TransformerLayer  # noqa: F401
