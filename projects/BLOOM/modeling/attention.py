from projects.BLOOM.modeling.bloom_model import BloomAttention  # noqa: F401
