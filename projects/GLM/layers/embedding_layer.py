"""Import-path shim: the reference keeps this class in its own file (projects/GLM/layers/embedding_layer.py); the implementation lives in projects/GLM/modeling_glm.py."""
from projects.GLM.modeling_glm import GLMEmbedding  # noqa: F401
