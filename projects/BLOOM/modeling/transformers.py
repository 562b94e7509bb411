from projects.BLOOM.modeling.bloom_model import BloomBlock  # noqa: F401
