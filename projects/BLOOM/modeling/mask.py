from projects.BLOOM.modeling.bloom_model import alibi_slopes, build_alibi_tensor  # noqa: F401
