"""Import-path shim: the reference keeps this class in its own file (projects/GLM/layers/attention_layer.py); the implementation lives in projects/GLM/modeling_glm.py."""
from projects.GLM.modeling_glm import GLMAttention as MultiheadAttention  # noqa: F401
