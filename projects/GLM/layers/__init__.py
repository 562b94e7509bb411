from projects.GLM.modeling_glm import GLMAttention as MultiheadAttention, GLMEmbedding, GLMLayer as TransformerLayer  # noqa: F401
